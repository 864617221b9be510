"""Pins the CPU oracle's preprocessing maps against the hand-derived golden vectors
(reference src/preprocess.jl:24-169) and against structural
invariants on the real Gset G1 fixture."""
import numpy as np
import pytest

from helpers import MAP_KEYS, g1_graph, k2_graph, load_golden, p3_graph


def _maps(sp, oracle_mod, C, As, bs):
    data = sp.SDPData(C, As, bs)
    asm = sp.assemble_sparse(data)
    o = oracle_mod.Oracle(asm, data.b)
    return data, asm, o, o.pattern_export()


@pytest.mark.parametrize("name,graph,fam", [("k2_maxcut.json", k2_graph, "maxcut"), ("p3_lovasz.json", p3_graph, "lovasz_theta")])
def test_golden_maps(sp, oracle_mod, name, graph, fam):
    gold = load_golden(name)
    C, As, bs = getattr(sp.problems, fam)(graph())
    data, asm, o, maps = _maps(sp, oracle_mod, C, As, bs)
    assert (data.n, data.m, len(asm.gids)) == (gold["n"], gold["m"], gold["nA"])
    assert o.pattern_sizes() == (gold["nnzT"], gold["nnzF"], gold["Ec"])
    for k in MAP_KEYS:
        np.testing.assert_array_equal(maps[k], np.array(gold[k]), err_msg=k)  # bit-exact, floats included


def test_k2_operator_identities(sp, oracle_mod):
    """SURVEY Appendix C 'Check': A(RR') = [a^2, b^2, -.25a^2 + .5ab - .25b^2] and S entries."""
    C, As, bs = sp.problems.maxcut(k2_graph())
    data, asm, o, _ = _maps(sp, oracle_mod, C, As, bs)
    a, b = 0.7, -1.3
    out = o.A_uu(np.array([[a], [b]]))
    np.testing.assert_allclose(out, [a * a, b * b, -.25 * a * a + .5 * a * b - .25 * b * b], rtol=0, atol=1e-15)
    y = np.array([0.3, -0.9, 1.0])
    o.At_preprocess(y)
    np.testing.assert_allclose(o.view("triuS", 3), [y[0] - .25, .25, y[1] - .25], atol=1e-16)
    np.testing.assert_allclose(o.view("S", 4), [y[0] - .25, .25, .25, y[1] - .25], atol=1e-16)


def test_p3_operator_identity(sp, oracle_mod):
    C, As, bs = sp.problems.lovasz_theta(p3_graph())
    data, asm, o, _ = _maps(sp, oracle_mod, C, As, bs)
    R = np.random.default_rng(3).standard_normal((3, 2))
    out = o.A_uu(R)
    exp = [2 * R[0] @ R[1], 2 * R[1] @ R[2], np.sum(R * R), -np.sum(R.sum(axis=0) ** 2)]
    np.testing.assert_allclose(out, exp, rtol=0, atol=1e-13)


def test_g1_pattern_counts(sp, oracle_mod):
    """G1 facts: n=800, nnzT = E+n = 19976, nnzF = 39152, E_c = 20776."""
    C, As, bs = sp.problems.maxcut(g1_graph())
    data, asm, o, maps = _maps(sp, oracle_mod, C, As, bs)
    assert (data.n, data.m) == (800, 800)
    assert o.pattern_sizes() == (19976, 39152, 20776)
    # structural invariants of the maps (Appendix B)
    tc, tr, fc, fr, mp = (maps[k] for k in ["triu_colptr", "triu_rowval", "full_colptr", "full_rowval", "mapped"])
    assert tc[0] == 1 and tc[-1] == 19977 and fc[-1] == 39153
    for col in range(800):
        rows = tr[tc[col] - 1: tc[col + 1] - 1]
        assert np.all(np.diff(rows) > 0) and np.all(rows <= col + 1)
    cols_full = np.repeat(np.arange(1, 801), np.diff(fc))
    lo, hi = np.minimum(fr, cols_full), np.maximum(fr, cols_full)
    tcol = np.repeat(np.arange(1, 801), np.diff(tc))
    assert np.array_equal(tr[mp - 1], lo) and np.array_equal(tcol[mp - 1], hi)
    # entries: first 800 are the diagonal constraints, then triu(C) in CSC order
    assert np.array_equal(maps["matptr"][:801], np.arange(1, 802)) and maps["matptr"][-1] == 20777
    assert np.array_equal(np.sort(maps["nzind"][800:]), np.arange(1, 19977))


def test_asymmetric_storage_flagged(sp, oracle_mod):
    """A lower entry with no mirrored upper entry leaves mappedto_triu = 0 (Appendix A.10)."""
    import scipy.sparse as sps
    n = 3
    Cm = sps.csc_matrix(np.eye(n))
    bad = sp.SparseMatrixCOO([2], [0], [1.0], n)  # only (3,1), no (1,3)
    data = sp.SDPData(Cm, [bad], np.zeros(1))
    o = oracle_mod.Oracle(sp.assemble_sparse(data), data.b)
    assert o.preprocess_rc == 1
    maps = o.pattern_export()
    assert (maps["mapped"] == 0).sum() == 1
