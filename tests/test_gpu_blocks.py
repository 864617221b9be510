"""Structured problem blocks (sdplrp_preprocess_blocks): the constraint families of the reference's problem
constructors (test/problem.jl:16-30 MaxCut, :50-62 Lovasz theta, :80-92 minimum bisection, :100-110 cut-norm) handed over as
DIAG / EDGES / IDENTITY / CSC descriptors must give the nine preprocessing maps of the triplet path bit for bit -- on the
shapes of the BASELINE configs C1-C4 -- and the same f!/g! values."""
import numpy as np
import pytest

from helpers import MAP_KEYS


def _cases(sp):
    P = sp.problems
    rng = np.random.default_rng(4)
    import scipy.sparse as sps
    A4 = sps.random(500, 500, density=0.02, random_state=np.random.RandomState(4), data_rvs=rng.standard_normal, format="csc")
    return [
        ("C1 maxcut", P.maxcut(P.gnm_graph(800, 19176, 1))),
        ("C2 lovasz", P.lovasz_theta(P.erdos_renyi(2000, 0.01, 2))),
        ("C3 bisection", P.minimum_bisection(P.erdos_renyi(20000, 10.0 / 20000, 3))),
        ("C4 cutnorm", P.cutnorm(A4)),
    ]


def test_block_descriptors_host():
    """CPU: the translator picks the structured kinds (one descriptor per family, not one per matrix)."""
    import sdplrplus.jl_b200 as sp
    from sdplrplus.jl_b200.types import structured_blocks
    kinds = {}
    for name, (C, As, bs) in _cases(sp):
        blocks, lowrank = structured_blocks(sp.SDPData(C, As, bs))
        kinds[name] = [b["kind"] for b in blocks]
    L = sp._lib
    assert kinds["C1 maxcut"] == [L.BLOCK_DIAG, L.BLOCK_CSC]
    assert kinds["C2 lovasz"] == [L.BLOCK_EDGES, L.BLOCK_IDENTITY]          # C = -11' is low rank
    assert kinds["C3 bisection"] == [L.BLOCK_DIAG, L.BLOCK_CSC]             # + the low-rank 11' constraint
    assert kinds["C4 cutnorm"] == [L.BLOCK_DIAG, L.BLOCK_CSC]


@pytest.mark.gpu
@pytest.mark.parametrize("config", ["default", "relabel"])
def test_blocks_give_the_maps_of_the_triplet_path(sp, gpu_handle_factory, config):
    from sdplrplus.jl_b200.types import structured_blocks, assemble_sparse
    for name, (C, As, bs) in _cases(sp):
        data = sp.SDPData(C, As, bs)
        asm = assemble_sparse(data)
        h1 = gpu_handle_factory(config)
        h1.preprocess(asm.n, asm.m, asm.mat_off, asm.I, asm.J, asm.V, asm.gids)
        m1 = h1.pattern_export()
        blocks, lowrank = structured_blocks(data)
        h2 = gpu_handle_factory(config)
        h2.preprocess_blocks(data.n, data.m, blocks)
        m2 = h2.pattern_export()
        assert h1.pattern_sizes() == h2.pattern_sizes(), name
        for k in MAP_KEYS:
            assert np.array_equal(m1[k], m2[k]), (name, k)
        assert [g for g, _ in lowrank] == [g for g, _ in asm.lowrank]
        h1.close(); h2.close()


@pytest.mark.gpu
def test_blocks_from_device_arrays(sp, gpu_handle_factory):
    """on_device: descriptors whose arrays are CUDA tensors (a generator that builds the graph on the GPU)."""
    import torch
    from sdplrplus.jl_b200.types import structured_blocks
    C, As, bs = sp.problems.maxcut(sp.problems.gnm_graph(800, 19176, 1))
    data = sp.SDPData(C, As, bs)
    blocks, _ = structured_blocks(data)
    h_host = gpu_handle_factory("default"); h_host.preprocess_blocks(data.n, data.m, blocks)
    dev_blocks = []
    for b in blocks:
        d = dict(b)
        for k in ("I", "J", "V"):
            if k in d and d[k] is not None:
                d[k] = torch.from_numpy(np.ascontiguousarray(d[k])).cuda()
        dev_blocks.append(d)
    torch.cuda.synchronize()
    h_dev = gpu_handle_factory("default"); h_dev.preprocess_blocks(data.n, data.m, dev_blocks)
    ma, mb = h_host.pattern_export(), h_dev.pattern_export()
    for k in MAP_KEYS:
        assert np.array_equal(ma[k], mb[k]), k
    h_host.close(); h_dev.close()


def _expand_blocks_numpy(n, blocks, L):
    """numpy restatement of the device expansion of csrc/blocks.cu (findnz order per matrix)."""
    Is, Js, Vs, lens, gids = [], [], [], [], []
    for b in blocks:
        kind, g0 = b["kind"], b["first_gid"]
        if kind == L.BLOCK_TRIPLETS:
            Is.append(b["I"]); Js.append(b["J"]); Vs.append(b["V"]); lens.append([len(b["V"])]); gids.append([g0])
        elif kind == L.BLOCK_CSC:
            colptr = np.asarray(b["J"]) - 1
            cols = np.repeat(np.arange(1, n + 1), np.diff(colptr))
            Is.append(b["I"]); Js.append(cols); Vs.append(b["V"]); lens.append([len(b["V"])]); gids.append([g0])
        elif kind == L.BLOCK_DIAG:
            k = b["count"]
            pos = np.asarray(b["I"]) if b.get("I") is not None else np.arange(1, k + 1)
            val = np.asarray(b["V"]) if b.get("V") is not None else np.ones(k)
            Is.append(pos); Js.append(pos); Vs.append(val); lens.append(np.ones(k, np.int64)); gids.append(np.arange(g0, g0 + k))
        elif kind == L.BLOCK_EDGES:
            k = b["count"]
            u, v = np.asarray(b["I"]), np.asarray(b["J"])
            w = np.asarray(b["V"]) if b.get("V") is not None else np.ones(k)
            Is.append(np.stack([u, v], 1).reshape(-1)); Js.append(np.stack([v, u], 1).reshape(-1)); Vs.append(np.repeat(w, 2))
            lens.append(np.full(k, 2, np.int64)); gids.append(np.arange(g0, g0 + k))
        else:
            idx = np.arange(1, n + 1)
            Is.append(idx); Js.append(idx); Vs.append(np.ones(n)); lens.append([n]); gids.append([g0])
    cat = lambda xs, dt: np.concatenate([np.asarray(x, dt) for x in xs]) if xs else np.zeros(0, dt)
    return cat(Is, np.int64), cat(Js, np.int64), cat(Vs, np.float64), np.concatenate([[0], np.cumsum(cat(lens, np.int64))]), cat(gids, np.int64)


def test_block_expansion_spec_equals_the_triplet_assembly():
    """CPU: the expansion rule of the structured blocks (restated in numpy) reproduces assemble_sparse's triplet stream --
    entry order inside every matrix, matrix offsets and global ids -- for the four families."""
    import sdplrplus.jl_b200 as sp
    from sdplrplus.jl_b200.types import structured_blocks, assemble_sparse
    for name, (C, As, bs) in _cases(sp):
        data = sp.SDPData(C, As, bs)
        asm = assemble_sparse(data)
        blocks, _ = structured_blocks(data)
        I, J, V, off, gids = _expand_blocks_numpy(data.n, blocks, sp._lib)
        assert np.array_equal(I, asm.I) and np.array_equal(J, asm.J) and np.array_equal(V, asm.V), name
        assert np.array_equal(off, asm.mat_off) and np.array_equal(gids, asm.gids), name
