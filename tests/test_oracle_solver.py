"""Known-answer tests of the reference (test/maxcut.jl, test/minimumbisection.jl) run
through the host driver `_sdplr` on the CPU oracle, plus the real Gset G1 instance."""
import numpy as np
import pytest

from helpers import g1_graph, k2_graph


def _solve(sp, oracle_mod, C, As, bs, r, **kw):
    return sp.sdplr(C, As, bs, r, engine_factory=oracle_mod.OracleEngine, printlevel=0, **kw)


def test_k2_maxcut(sp, oracle_mod):
    C, As, bs = sp.problems.maxcut(k2_graph())
    res = _solve(sp, oracle_mod, C, As, bs, 1, fprec=0.0, gtol=1e-8, objtol=1e-8, ptol=1e-8, prior_trace_bound=2.0)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7)                 # test/maxcut.jl:24 (isapprox default rtol ~1.5e-8)


def test_k2_maxcut_sigma0(sp, oracle_mod):
    C, As, bs = sp.problems.maxcut(k2_graph())
    res = _solve(sp, oracle_mod, C, As, bs, 1, sigma_0=10.0, fprec=0.0, gtol=1e-8, objtol=1e-8, ptol=1e-8, prior_trace_bound=2.0)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7)                 # test/maxcut.jl:47


def test_k2_maxcut_init_func(sp, oracle_mod):
    C, As, bs = sp.problems.maxcut(k2_graph())
    rng = np.random.default_rng(5)

    def init_func(data, r, sigma):
        return rng.standard_normal((data.n, r)) * np.sqrt(sigma), np.zeros(data.m)

    res = _solve(sp, oracle_mod, C, As, bs, 1, init_func=init_func, init_args=(10.0,), fprec=0.0, gtol=1e-8, objtol=1e-8,
                 ptol=1e-8, prior_trace_bound=2.0)
    assert res["obj"] == pytest.approx(-1.0, rel=1e-7)                 # test/maxcut.jl:75


def test_k2_minimum_bisection(sp, oracle_mod):
    C, As, bs = sp.problems.minimum_bisection(k2_graph())
    res = _solve(sp, oracle_mod, C, As, bs, 1, fprec=0.0, objtol=1e-4, ptol=1e-4, prior_trace_bound=2.0)
    assert (res["obj"] - 1) / (1 + abs(res["obj"])) < 1e-4             # test/minimumbisection.jl:22


def test_g1_maxcut_default_tolerances(sp, oracle_mod):
    """Protocol of exps/batch_test.txt: rank 10, ptol = objtol = 1e-2, trace bound n."""
    C, As, bs = sp.problems.maxcut(g1_graph())
    res = _solve(sp, oracle_mod, C, As, bs, 10, prior_trace_bound=800.0, seed=0)
    assert res["primal_vio"] <= 1e-2 and res["min_duality_gap"] <= 1e-2
    # G1's SDP optimum is 12083.2 (Gset literature); the solve is at 1e-2 tolerance
    assert abs(-res["obj"] - 12083.2) / 12083.2 < 1e-2
    assert res["max_dual_value"] <= res["obj"] + 1e-9 * abs(res["obj"]) or res["min_duality_gap"] < 0
    assert 50 < res["iter"] < 2000 and res["majoriter"] < 30


def test_sub_solver_hook(sp, oracle_mod):
    """SDPLRPlus.Solver / SolverCore.solve! (src/lowrankopt.jl:4-53): keywords applied to the config (unknown ones reported
    and skipped), flat start vector, solution / multipliers / elapsed_time / status filled."""
    from types import SimpleNamespace
    C, As, bs = sp.problems.maxcut(k2_graph())
    model = SimpleNamespace(C=C, As=As, b=bs, rank=1)
    solver = sp.Solver(model, engine_factory=oracle_mod.OracleEngine, printlevel=0, fprec=0.0, gtol=1e-8, objtol=1e-8, ptol=1e-8,
                       prior_trace_bound=2.0, no_such_option=3)
    assert solver.config.ptol == 1e-8 and solver.Rt0.shape == (2,) and solver.lambda0.shape == (2,)
    stats = solver.solve(model, sp.GenericExecutionStats(), maxtime=60.0, another_unknown=1)
    assert solver.config.maxtime == 60.0
    assert stats.status == "first_order" and stats.elapsed_time > 0
    assert stats.solution.shape == (2,) and stats.multipliers.shape == (2,)
    assert stats.objective == pytest.approx(-1.0, rel=1e-7)                       # K2 MaxCut, test/maxcut.jl:24
    assert abs(abs(stats.solution[0]) - 1.0) < 1e-6 and stats.solution[0] * stats.solution[1] < 0   # the cut: R = (+-1, -+1)
    C, As, bs = sp.problems.maxcut(g1_graph())
    st = sp.Solver(SimpleNamespace(C=C, As=As, b=bs, rank=10), engine_factory=oracle_mod.OracleEngine, printlevel=0,
                   prior_trace_bound=800.0).solve()
    assert st.solution.shape == (8000,) and st.solver_specific["primal_vio"] <= 1e-2 and st.solver_specific["min_duality_gap"] <= 1e-2
