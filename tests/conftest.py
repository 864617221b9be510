import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sp():
    import sdplrplus.jl_b200 as sp_
    return sp_


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


# kernel configurations every GPU parity test runs under: the product default (auto relabeling), forced hub-first relabeling
# (exercises every permuting copy on the small regular test graphs too), the literal vector two-loop, the two-phase (hub |
# tail columns) gather pass with a 5-column hub prefix (both phases and the accumulate-with-dots epilogue are non-trivial on
# the small graphs; the multi-GPU overlap runs on the same machinery), and the asynchronous tile pipeline of gather.cu
# (cp.async.bulk and cp.async row gathers; small tiles so that chunked long rows, split rows and multi-pass tiles all occur)
GPU_CONFIGS = {"default": {}, "relabel": {"relabel": 1}, "literal": {"relabel": 1, "lbfgs_kernel": 0},
               "phases": {"relabel": 1, "spmm_phases": 5},
               "gather_bulk": {"relabel": 1, "gather_mode": 1, "gather_tile": 16},
               "gather_async": {"gather_mode": 2},
               "gather_async_small": {"relabel": 1, "gather_mode": 2, "gather_tile": 24, "gather_stages": 3, "gather_hints": 1}}
GPU_CONFIG_PARAMS = list(GPU_CONFIGS)


@pytest.fixture(scope="session")
def gpu_handle_factory(sp):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")

    def make(config="default"):
        h = sp.Handle(device=0)
        for k, v in GPU_CONFIGS[config].items():
            h.set_option(k, v)
        return h
    return make
