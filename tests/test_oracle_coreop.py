"""The reference's own operator tests (test/coreop.jl) run against the CPU oracle:
dense identities at 1e-10, finite-difference gradient at 1e-8, incremental residual
recurrence after linesearch!, inequality capping, and the adjoint operator."""
import numpy as np
import pytest

from helpers import COMBOS, dense_S, dense_primal_vio, families, make_case


def _engine(sp, oracle_mod, data, Rt0, r, sigma=2.0, h=4):
    eng = oracle_mod.OracleEngine(data)
    eng.init_vars(r, Rt0, np.zeros(data.m), sigma, h)
    return eng


def _fd_gradient(eng, Rt0, n, r):
    """central differences of f! (test/coreop.jl:19-32)"""
    g = np.zeros(n * r)
    x = Rt0.reshape(-1).copy()
    hstep = 6e-6
    for k in range(x.size):
        xp = x.copy(); xp[k] += hstep
        eng.set_R(xp); fp, _ = eng.f()
        xm = x.copy(); xm[k] -= hstep
        eng.set_R(xm); fm, _ = eng.f()
        g[k] = (fp - fm) / (2 * hstep)
    eng.set_R(x)
    return g


FAMS = ["maxcut", "lovasz_theta", "minimum_bisection", "cutnorm", "mu_conductance_0.01", "mu_conductance_0.05", "mu_conductance_0.1"]


@pytest.mark.parametrize("fam", FAMS)
@pytest.mark.parametrize("seed,n,p,r", COMBOS)
def test_f_g_linesearch(sp, oracle_mod, fam, seed, n, p, r):
    fn = dict(families(sp))[fam]
    data, Rt0, rng = make_case(sp, fn, seed, n, p, r)
    eng = _engine(sp, oracle_mod, data, Rt0, r)
    eng.f()
    assert np.max(np.abs(eng.get_pvio_raw() - dense_primal_vio(data, Rt0))) < 1e-10      # test/coreop.jl:58-61
    gnum = _fd_gradient(eng, Rt0, data.n, r)
    eng.f(); eng.g()
    gana = eng.get_G().reshape(-1)
    assert np.max(np.abs(gnum - gana)) / (1 + np.max(np.abs(gana))) < 1e-8               # :19-32
    eng.set_D(-eng.get_G())
    bq = eng.linesearch_coeffs()
    alpha, _ = sp.pick_alpha(bq, 1.0)
    eng.step(alpha)
    Rnew = Rt0 + alpha * (-gana.reshape(data.n, r))
    np.testing.assert_allclose(eng.get_R(), Rnew, atol=1e-14)
    assert np.max(np.abs(eng.get_pvio_raw() - dense_primal_vio(data, Rnew))) < 1e-10     # :65-72
    # oracle's own root finder agrees with the host one
    a2 = np.zeros(1); f2 = np.zeros(1)
    import ctypes as C
    rc = oracle_mod.load().orc_pick_alpha(bq.ctypes.data_as(C.POINTER(C.c_double)), 1.0,
                                          a2.ctypes.data_as(C.POINTER(C.c_double)), f2.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0 and abs(a2[0] - alpha) < 1e-9 * max(1.0, abs(alpha))


@pytest.mark.parametrize("mu", [0.01, 0.05, 0.1])
@pytest.mark.parametrize("seed,n,p,r", COMBOS)
def test_inequalities(sp, oracle_mod, mu, seed, n, p, r):
    data, Rt0, rng = make_case(sp, lambda A: sp.problems.mu_conductance_ineq(A, mu), seed, n, p, r)
    eng = _engine(sp, oracle_mod, data, Rt0, r)
    eng.f()
    pv = dense_primal_vio(data, Rt0)
    assert np.max(np.abs(eng.get_pvio_raw() - pv)) < 1e-10
    cap = pv[: data.m].copy()
    cap[data.constraint_types] = np.maximum(cap[data.constraint_types], 0.0)
    assert np.max(np.abs(eng.o.view("pvio") - cap)) < 1e-10                               # test/coreop.jl:107-112
    gnum = _fd_gradient(eng, Rt0, data.n, r)
    eng.f(); eng.g()
    gana = eng.get_G().reshape(-1)
    assert np.max(np.abs(gnum - gana)) / (1 + np.max(np.abs(gana))) < 1e-8


@pytest.mark.parametrize("fam", ["maxcut", "lovasz_theta", "minimum_bisection", "mu_conductance_0.01", "mu_conductance_0.05",
                                 "mu_conductance_0.1", "ineq_0.01", "ineq_0.05", "ineq_0.1"])
@pytest.mark.parametrize("seed,n,p,r", COMBOS)
def test_adjoint(sp, oracle_mod, fam, seed, n, p, r):
    if fam.startswith("ineq_"):
        mu = float(fam.split("_")[1])
        fn = lambda A: sp.problems.mu_conductance_ineq(A, mu)
    else:
        fn = dict(families(sp))[fam]
    data, Rt0, rng = make_case(sp, fn, seed, n, p, r)
    eng = _engine(sp, oracle_mod, data, Rt0, r)
    eng.f()
    y = np.random.default_rng(seed + 100).standard_normal(data.m + 1)
    eng.o.At_preprocess(y)
    S = dense_S(data, y)
    Yl = eng.o.At_left(Rt0)                                                               # test/coreop.jl:160-165
    assert np.max(np.abs(Yl - (Rt0.T @ S).T)) < 1e-10
    x = rng.standard_normal((data.n, r))
    assert np.max(np.abs(eng.o.At_right(x) - S @ x)) < 1e-10                              # :167-172
