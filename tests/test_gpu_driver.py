"""The native host driver (sdplrp_solve, csrc/driver.cu: _sdplr of src/sdplr.jl:140-449 inside the library) against
the Python mirror of the same loop driving the same entry points one by one, and against the oracle-driven solve."""
import numpy as np
import pytest

from helpers import g1_graph, k2_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["default", "relabel"])
def handle(gpu_handle_factory, request):
    h = gpu_handle_factory(request.param)
    yield h
    h.close()


def _problem(sp, which):
    P = sp.problems
    if which == "g1_maxcut":
        return P.maxcut(g1_graph()) + (None,), 10, dict(prior_trace_bound=800.0)
    if which == "lovasz":   # the 5-cycle, theta = sqrt(5) (Erdos-Renyi theta instances need 1e4+ iterations: not a unit test)
        import scipy.sparse as sps
        i5 = np.arange(5)
        G = sps.csc_matrix((np.ones(10), (np.r_[i5, (i5 + 1) % 5], np.r_[(i5 + 1) % 5, i5])), shape=(5, 5))
        return P.lovasz_theta(G) + (None,), 3, dict(prior_trace_bound=1.0)
    if which == "bisect":
        return P.minimum_bisection(P.erdos_renyi(200, 0.05, 4)) + (None,), 8, dict(prior_trace_bound=200.0)
    if which == "cutnorm":
        import scipy.sparse as sps
        g = np.random.default_rng(4)
        A = sps.csc_matrix(g.standard_normal((80, 80)) * (g.random((80, 80)) < 0.1))
        return P.cutnorm(A) + (None,), 8, dict(prior_trace_bound=160.0)
    if which == "ineq":
        return P.mu_conductance_ineq(P.erdos_renyi(60, 0.2, 5), 0.05), 6, dict(prior_trace_bound=1.0, maxiter=3000)
    raise KeyError(which)


@pytest.mark.parametrize("which", ["g1_maxcut", "lovasz", "bisect", "cutnorm", "ineq"])
def test_native_driver_matches_python_driver(sp, handle, which):
    (C, As, bs, types), r, kw = _problem(sp, which)
    fac = lambda data: sp.B200Engine(data, handle=handle)
    kw = dict(kw, printlevel=0, seed=3, rng_stream="native", maxtime=30.0)
    kw.setdefault("maxiter", 20000)   # bounded: a wrong decision must fail, not spin
    rp = sp.sdplr(C, As, bs, r, types, engine_factory=fac, driver="python", **kw)
    rn = sp.sdplr(C, As, bs, r, types, engine_factory=fac, driver="native", **kw)
    # same entry points, same order, same random numbers: the two loops take the same decisions; the step sizes differ
    # in their last bits (numpy.roots vs the bracketing solver), which the iteration amplifies to ~1e-8
    assert rn["iter"] == rp["iter"] and rn["majoriter"] == rp["majoriter"] and rn["r"] == rp["r"]
    assert rn["lanczos_steps"] == rp["lanczos_steps"]
    for k in ("obj", "max_dual_value", "primal_vio", "grad_norm", "sigma", "L"):
        assert rn[k] == pytest.approx(rp[k], rel=1e-6, abs=1e-9), k
    np.testing.assert_allclose(rn["Rt"], rp["Rt"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rn["lambda"][: len(bs)], np.asarray(rp["lambda"])[: len(bs)], rtol=1e-5, atol=1e-6)
    assert rn["status"] == 0 or which == "ineq"
    if which == "lovasz":
        assert abs(rn["obj"] + np.sqrt(5.0)) <= 2e-2 * np.sqrt(5.0)


def test_native_driver_k2_known_answers(sp, handle):
    fac = lambda data: sp.B200Engine(data, handle=handle)
    C, As, bs = sp.problems.maxcut(k2_graph())
    kw = dict(engine_factory=fac, driver="native", printlevel=0, fprec=0.0, gtol=1e-8, objtol=1e-8, ptol=1e-8, prior_trace_bound=2.0,
              maxtime=60.0)
    assert sp.sdplr(C, As, bs, 1, **kw)["obj"] == pytest.approx(-1.0, rel=1e-7)                    # test/maxcut.jl:24
    assert sp.sdplr(C, As, bs, 1, sigma_0=10.0, **kw)["obj"] == pytest.approx(-1.0, rel=1e-7)      # test/maxcut.jl:47
    res = sp.sdplr(C, As, bs, 1, eigval_highprecision=True, eval_DIMACS_errs=True, **kw)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7) and np.abs(res["DIMACS_errs"]).max() < 1e-6
    C, As, bs = sp.problems.minimum_bisection(k2_graph())
    res = sp.sdplr(C, As, bs, 1, engine_factory=fac, driver="native", printlevel=0, fprec=0.0, objtol=1e-4, ptol=1e-4,
                   prior_trace_bound=2.0, maxtime=60.0)
    assert (res["obj"] - 1) / (1 + abs(res["obj"])) < 1e-4                                         # test/minimumbisection.jl:22


def test_native_driver_matches_oracle_solve(sp, oracle_mod, handle):
    """whole solve on G1, native driver on the GPU vs the Python loop on the CPU oracle: objective within 1e-6"""
    C, As, bs = sp.problems.maxcut(g1_graph())
    kw = dict(printlevel=0, prior_trace_bound=800.0, seed=0, objtol=float("inf"), maxtime=120.0)
    rn = sp.sdplr(C, As, bs, 10, engine_factory=lambda d: sp.B200Engine(d, handle=handle), driver="native", **kw)
    ro = sp.sdplr(C, As, bs, 10, engine_factory=oracle_mod.OracleEngine, **kw)
    assert rn["iter"] == ro["iter"] and rn["majoriter"] == ro["majoriter"]
    assert rn["obj"] == pytest.approx(ro["obj"], rel=1e-6)
    assert rn["primal_vio"] <= 1e-2


def test_native_driver_rank_update_and_budget(sp, handle):
    """rank_update! (src/coreop.jl:518-526): a rank-1 start cannot close the gap of a MaxCut SDP; the driver doubles r
    with a fresh random point from the device generator.  And: an exhausted iteration budget reports status 1."""
    P = sp.problems
    C, As, bs = P.maxcut(P.erdos_renyi(80, 0.15, 6))
    fac = lambda data: sp.B200Engine(data, handle=handle)
    res = sp.sdplr(C, As, bs, 1, engine_factory=fac, driver="native", printlevel=0, prior_trace_bound=80.0, rankupd_tol=1,
                   objtol=1e-3, ptol=1e-3, seed=5, maxtime=120.0)
    assert res["r"] >= 2 and res["status"] == 0
    assert res["primal_vio"] <= 1e-3 and res["min_duality_gap"] <= 1e-3
    assert res["Rt"].shape == (80, res["r"])
    res = sp.sdplr(C, As, bs, 4, engine_factory=fac, driver="native", printlevel=0, prior_trace_bound=80.0, maxiter=5, seed=5)
    assert res["status"] == 1 and 5 <= res["iter"] <= 7


def test_fill_uniform_is_layout_independent(sp, gpu_handle_factory):
    """the device generator is keyed by the reference vertex: relabeling must not change the matrix the host sees"""
    P = sp.problems
    C, As, bs = P.maxcut(P.powerlaw_graph(3000, 20000, 1))
    data = sp.SDPData(C, As, bs)
    out = []
    for cfg in ("default", "relabel"):
        h = gpu_handle_factory(cfg)
        eng = sp.B200Engine(data, handle=h)
        h.set_rank(7, 2)
        h.fill_uniform(sp._lib.MAT_R, 1234)
        out.append(eng.get_R())
        h.close()
    np.testing.assert_array_equal(out[0], out[1])
    R = out[0]
    assert R.shape == (3000, 7) and -1.0 <= R.min() < -0.99 and 0.99 < R.max() < 1.0
    assert abs(R.mean()) < 0.02 and abs(R.var() - 1.0 / 3.0) < 0.02
    assert np.unique(R).size == R.size


def test_iterate_matches_python_loop(sp, handle):
    """sdplrp_iterate == run_inner_iterations driven call by call (what bench.py times)"""
    C, As, bs = sp.problems.maxcut(g1_graph())
    data = sp.SDPData(C, As, bs)
    Rt0 = 2 * np.random.default_rng(0).random((data.n, 10)) - 1
    outs = []
    for native in (False, True):
        eng = sp.B200Engine(data, handle=handle)
        eng.init_vars(10, Rt0, np.zeros(data.m), 2.0, 4)
        eng.fg()
        last = sp.solver.run_inner_iterations(eng, 25, native=native)
        outs.append((np.array(last), eng.get_R()))
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(outs[1][1], outs[0][1], rtol=1e-7, atol=1e-9)
