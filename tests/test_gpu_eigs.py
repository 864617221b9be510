"""GPU parity of the high-precision eigenvalue path and the DIMACS errors, through the C ABI:
sdplrp_S_eigval (thick-restart Lanczos on the device) against dense eigenvalues and the ARPACK
oracle, sdplrp_dual_obj_highprecision and sdplrp_dimacs_errors against the oracle
(SDP_S_eigval / dual_obj / DIMACS_errors, src/coreop.jl:351-453)."""
import numpy as np
import pytest

from helpers import COMBOS, dense_S, families, g1_graph, k2_graph, make_case, p3_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["default", "relabel"])
def handle(gpu_handle_factory, request):
    h = gpu_handle_factory(request.param)
    yield h
    h.close()


def _pair(sp, oracle_mod, handle, data, Rt0, r, lam0):
    ge = sp.B200Engine(data, handle=handle)
    ge.init_vars(r, Rt0, lam0, 2.0, 4)
    oe = oracle_mod.OracleEngine(data)
    oe.init_vars(r, Rt0, lam0, 2.0, 4)
    return ge, oe


@pytest.mark.parametrize("fam", ["maxcut", "lovasz_theta", "minimum_bisection", "cutnorm", "mu_conductance_0.05"])
@pytest.mark.parametrize("seed,n,p,r", COMBOS[::2])
def test_S_eigval_small_shapes(sp, oracle_mod, handle, fam, seed, n, p, r):
    """n = 5..25 (ncv = n: the Krylov space is exhausted, the Ritz values are exact)"""
    data, Rt0, rng = make_case(sp, dict(families(sp))[fam], seed, n, p, r)
    ge, oe = _pair(sp, oracle_mod, handle, data, Rt0, r, rng.standard_normal(data.m))
    ge.fg(); oe.fg()
    v0 = rng.standard_normal(data.n)
    dg, eg, mv = ge.dual_obj_highprecision(float(data.n), v0)
    do, eo, _ = oe.dual_obj_highprecision(float(data.n))
    lam = np.linalg.eigvalsh(dense_S(data, oe.get_y()))
    assert abs(eg - lam[0]) <= 1e-9 * max(1.0, abs(lam).max())
    assert abs(dg - do) <= 1e-6 * max(1.0, abs(do))
    assert 1 <= mv <= 4 * data.n
    ev, bd, mv, rs = handle.S_eigval(nevs=min(3, data.n - 1), tol=0.0, v0=v0)
    np.testing.assert_allclose(ev, lam[: ev.size], rtol=0, atol=1e-9 * max(1.0, abs(lam).max()))


@pytest.mark.parametrize("graph,fam,n", [(k2_graph, "maxcut", 2), (p3_graph, "lovasz_theta", 3), (k2_graph, "minimum_bisection", 2)])
def test_S_eigval_tiny(sp, oracle_mod, handle, graph, fam, n):
    C, As, bs = getattr(sp.problems, fam)(graph())
    data = sp.SDPData(C, As, bs)
    Rt0 = 2 * np.random.default_rng(3).random((data.n, 1)) - 1
    ge, oe = _pair(sp, oracle_mod, handle, data, Rt0, 1, np.linspace(-0.3, 0.4, data.m))
    ge.fg(); oe.fg()
    dg, eg, _ = ge.dual_obj_highprecision(2.0, np.arange(1.0, n + 1))
    do, eo, _ = oe.dual_obj_highprecision(2.0)
    assert eg == pytest.approx(eo, abs=1e-10) and dg == pytest.approx(do, abs=1e-10)


@pytest.mark.parametrize("fam", ["maxcut", "lovasz_theta", "minimum_bisection", "cutnorm"])
def test_S_eigval_restarts(sp, oracle_mod, handle, fam):
    """n = 400 > ncv: several thick restarts; host and device start vectors; reference tolerances 1e-6 and 0"""
    P = sp.problems
    if fam == "cutnorm":
        import scipy.sparse as sps
        g = np.random.default_rng(4)
        C, As, bs = P.cutnorm(sps.csc_matrix(g.standard_normal((200, 200)) * (g.random((200, 200)) < 0.05)))
    else:
        C, As, bs = getattr(P, fam)(P.erdos_renyi(400, 0.03, 8))
    data = sp.SDPData(C, As, bs)
    r = 6
    rng = np.random.default_rng(0)
    Rt0 = 2 * rng.random((data.n, r)) - 1
    ge, oe = _pair(sp, oracle_mod, handle, data, Rt0, r, 0.1 * rng.standard_normal(data.m))
    ge.fg(); oe.fg()
    v0 = rng.standard_normal(data.n)
    dg, eg, mv = ge.dual_obj_highprecision(float(data.n), v0)
    do, eo, _ = oe.dual_obj_highprecision(float(data.n))
    lam = np.linalg.eigvalsh(dense_S(data, oe.get_y()))
    scale = max(1.0, abs(lam[0] + 1.0))
    assert abs(eg - lam[0]) <= 1e-6 * scale and abs(eo - lam[0]) <= 1e-6 * scale   # both within the reference's tol of the truth
    assert abs(dg - do) <= 2e-6 * scale * data.n
    assert mv >= 100 or fam != "maxcut"
    # tol = 0 (machine precision, the DIMACS call), three eigenvalues, device-seeded start vector
    ev, bd, mv, rs = handle.S_eigval(nevs=3, ncv=40, tol=0.0, v0=None, seed=11)
    np.testing.assert_allclose(ev, lam[:3], rtol=0, atol=1e-9 * max(1.0, abs(lam).max()))
    assert rs >= 2 and np.all(bd <= 1e-9 * max(1.0, abs(lam).max()))
    # loose tolerance stops earlier
    ev2, _, mv2, _ = handle.S_eigval(nevs=1, ncv=40, tol=1e-3, v0=v0)
    assert mv2 <= mv and abs(ev2[0] - lam[0]) <= 1e-3 * scale


@pytest.mark.parametrize("fam", ["maxcut", "lovasz_theta", "minimum_bisection", "mu_conductance_0.05"])
def test_dimacs_errors_gpu(sp, oracle_mod, handle, fam):
    P = sp.problems
    out = dict(families(sp))[fam](P.erdos_renyi(300, 0.04, 9))
    data = sp.SDPData(*out)
    r = 5
    rng = np.random.default_rng(2)
    Rt0 = 2 * rng.random((data.n, r)) - 1
    ge, oe = _pair(sp, oracle_mod, handle, data, Rt0, r, 0.2 * rng.standard_normal(data.m))
    ge.fg(); oe.fg()
    normb = float(np.linalg.norm(data.b)); normC = sp.types.frobenius_norm(data.C)
    eg = ge.dimacs_errors(normb, normC, rng.standard_normal(data.n))
    eo = oe.dimacs_errors(normb, normC)
    for k in (0, 4, 5):
        assert abs(eg[k] - eo[k]) <= 1e-10 * max(1.0, abs(eo[k])), (k, eg, eo)
    assert eg[1] == 0.0 and eg[2] == 0.0
    assert abs(eg[3] - eo[3]) <= 1e-8 * max(1.0, abs(eo[3]))
    # the call leaves y = -lambda behind, as the reference does
    np.testing.assert_allclose(ge.get_y(), np.concatenate([-ge.get_lambda(), [1.0]]), rtol=0, atol=0)


def test_highprecision_solve_gpu(sp, oracle_mod, handle):
    """eigval_highprecision + eval_DIMACS_errs through the whole driver: same iterates as the oracle-driven solve"""
    C, As, bs = sp.problems.maxcut(k2_graph())
    fac = lambda data: sp.B200Engine(data, handle=handle)
    res = sp.sdplr(C, As, bs, 1, engine_factory=fac, printlevel=0, fprec=0.0, gtol=1e-8, objtol=1e-8, ptol=1e-8,
                   prior_trace_bound=2.0, eigval_highprecision=True, eval_DIMACS_errs=True, maxtime=60.0)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7)
    assert np.abs(res["DIMACS_errs"]).max() < 1e-6
    C, As, bs = sp.problems.maxcut(g1_graph())
    kw = dict(printlevel=0, prior_trace_bound=800.0, seed=0, eigval_highprecision=True, eval_DIMACS_errs=True, maxtime=120.0)
    rg = sp.sdplr(C, As, bs, 10, engine_factory=fac, **kw)
    ro = sp.sdplr(C, As, bs, 10, engine_factory=oracle_mod.OracleEngine, **kw)
    assert rg["iter"] == ro["iter"] and rg["majoriter"] == ro["majoriter"]
    assert rg["obj"] == pytest.approx(ro["obj"], rel=1e-6)
    assert rg["max_dual_value"] == pytest.approx(ro["max_dual_value"], rel=1e-5)
    np.testing.assert_allclose(rg["DIMACS_errs"], ro["DIMACS_errs"], rtol=1e-4, atol=1e-7)
