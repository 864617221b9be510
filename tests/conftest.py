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


@pytest.fixture(scope="session")
def gpu_handle_factory(sp):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return lambda: sp.Handle(device=0)
