"""CPU checks of the high-precision eigenvalue path and the DIMACS errors on the oracle
(SDP_S_eigval, dual_obj(highprecision=true), DIMACS_errors: src/coreop.jl:351-453), against dense
linear algebra, plus the host-side dense eigen-solver of the ABI (the projected problem of the
restarted Lanczos)."""
import numpy as np
import pytest

from helpers import COMBOS, dense_S, dense_of, families, g1_graph, k2_graph, make_case


@pytest.mark.parametrize("k", [1, 2, 3, 7, 40, 100])
def test_dense_symeig_host_helper(sp, k):
    rng = np.random.default_rng(k)
    A = rng.standard_normal((k, k)); A = A + A.T
    ev, Q = sp._lib.dense_symeig(A)
    scale = max(1.0, np.abs(A).max()) * k
    assert np.abs(ev - np.linalg.eigvalsh(A)).max() <= 1e-14 * scale
    assert np.abs(A @ Q - Q * ev).max() <= 1e-14 * scale
    assert np.abs(Q.T @ Q - np.eye(k)).max() <= 1e-14 * k * 4


def test_dense_symeig_arrowhead(sp):
    """the shape of the projected matrix after a thick restart: diagonal block + arrow + tridiagonal tail"""
    rng = np.random.default_rng(0)
    k, m = 20, 60
    T = np.zeros((m, m))
    T[np.arange(k), np.arange(k)] = np.sort(rng.standard_normal(k))
    T[:k, k] = T[k, :k] = 1e-3 * rng.standard_normal(k)
    for j in range(k, m):
        T[j, j] = rng.standard_normal()
    for j in range(k, m - 1):
        T[j, j + 1] = T[j + 1, j] = abs(rng.standard_normal())
    ev, Q = sp._lib.dense_symeig(T)
    assert np.abs(ev - np.linalg.eigvalsh(T)).max() <= 1e-12
    assert np.abs(T @ Q - Q * ev).max() <= 1e-12


@pytest.mark.parametrize("fam", ["maxcut", "lovasz_theta", "minimum_bisection", "cutnorm", "mu_conductance_0.05"])
@pytest.mark.parametrize("seed,n,p,r", COMBOS[2::3])
def test_oracle_S_eigval_and_dimacs_vs_dense(sp, oracle_mod, fam, seed, n, p, r):
    data, Rt0, rng = make_case(sp, dict(families(sp))[fam], seed, n, p, r)
    oe = oracle_mod.OracleEngine(data)
    lam0 = rng.standard_normal(data.m)
    oe.init_vars(r, Rt0, lam0, 2.0, 4)
    L, obj, _, _ = oe.fg()
    # high-precision dual value: y = -min(ub, lambda - sigma*raw), S(y), lambda_min by ARPACK vs dense
    d, lam, _ = oe.dual_obj_highprecision(float(data.n))
    y = oe.get_y()
    lam_dense = np.linalg.eigvalsh(dense_S(data, y))[0]
    assert abs(lam - lam_dense) <= 1e-6 * max(1.0, abs(lam_dense + 1.0))
    assert d == pytest.approx(-float(y[: data.m] @ data.b) + data.n * min(lam, 0.0), rel=1e-12, abs=1e-12)
    # DIMACS errors against the formulas of src/coreop.jl:417-425 evaluated densely
    normb = float(np.linalg.norm(data.b)); normC = sp.types.frobenius_norm(data.C)
    errs = oe.dimacs_errors(normb, normC)
    raw = oe.get_pvio_raw()
    lam_vec = oe.get_lambda()
    y2 = np.concatenate([-lam_vec, [1.0]])
    np.testing.assert_allclose(oe.get_y(), y2)
    # err6 uses the sparse part of S only (the reference's `var.Rt * aux.sparse_S`)
    S_sparse = np.zeros((data.n, data.n))
    Cd = data.C
    if not isinstance(Cd, sp.SymLowRankMatrix):
        S_sparse += dense_of(Cd)
    for i, A in enumerate(data.matrices()):
        if not isinstance(A, sp.SymLowRankMatrix):
            S_sparse += y2[i] * dense_of(A)
    S_full = dense_S(data, y2)
    X = Rt0 @ Rt0.T
    den = 1.0 + abs(obj) + abs(lam_vec @ data.b)
    want = [np.linalg.norm(raw[: data.m]) / (1 + normb), 0.0, 0.0, max(0.0, -np.linalg.eigvalsh(S_full)[0]) / (1 + normC),
            (obj - lam_vec @ data.b) / den, np.sum(X * S_sparse) / den]
    np.testing.assert_allclose(errs, want, rtol=1e-8, atol=1e-9)


def test_highprecision_solve_and_dimacs_on_oracle(sp, oracle_mod):
    """config.eigval_highprecision / eval_DIMACS_errs drive the same _sdplr loop (src/sdplr.jl:311-321, 419-425)."""
    C, As, bs = sp.problems.maxcut(k2_graph())
    res = sp.sdplr(C, As, bs, 1, engine_factory=oracle_mod.OracleEngine, printlevel=0, fprec=0.0, gtol=1e-8, objtol=1e-8,
                   ptol=1e-8, prior_trace_bound=2.0, eigval_highprecision=True, eval_DIMACS_errs=True)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7)
    e = res["DIMACS_errs"]
    assert e.shape == (6,) and e[1] == 0.0 and e[2] == 0.0
    assert abs(e[0]) < 1e-7 and abs(e[3]) < 1e-6 and abs(e[4]) < 1e-6 and abs(e[5]) < 1e-6
    C, As, bs = sp.problems.maxcut(g1_graph())
    res = sp.sdplr(C, As, bs, 10, engine_factory=oracle_mod.OracleEngine, printlevel=0, prior_trace_bound=800.0, seed=0,
                   eigval_highprecision=True, eval_DIMACS_errs=True)
    assert res["primal_vio"] <= 1e-2 and res["min_duality_gap"] <= 1e-2
    assert abs(-res["obj"] - 12083.2) / 12083.2 < 1e-2
    e = res["DIMACS_errs"]
    assert e[0] <= 1e-2 and 0 <= e[3] < 1e-2 and abs(e[4]) < 1e-2
