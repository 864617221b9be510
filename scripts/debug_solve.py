"""Debug helper: full solve on GPU vs oracle for one family; prints iteration counts and timings."""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from oracle import pyoracle
P = sp.problems
which = sys.argv[1] if len(sys.argv) > 1 else "lovasz"
kw = dict(printlevel=0, seed=0, maxtime=300.0, maxiter=200000)
types = None
if which == "lovasz":
    C, As, bs = P.lovasz_theta(P.erdos_renyi(200, 0.05, 2)); r = 10; kw["prior_trace_bound"] = 1.0
elif which == "bisect":
    C, As, bs = P.minimum_bisection(P.erdos_renyi(500, 0.02, 3)); r = 10; kw["prior_trace_bound"] = 500.0
else:
    C, As, bs, types = P.mu_conductance_ineq(P.erdos_renyi(60, 0.15, 5), 0.05); r = 5
    kw.update(objtol=math.inf, ptol=1e-3, maxmajoriter=40)
for name, fac in (("gpu", lambda d: sp.B200Engine(d, handle=sp.Handle(device=0))), ("oracle", pyoracle.OracleEngine)):
    t0 = time.perf_counter()
    res = sp.sdplr(C, As, bs, r, constraint_types=types, engine_factory=fac, **kw)
    print(name, {k: res[k] for k in ("iter", "majoriter", "obj", "primal_vio", "max_dual_value", "min_duality_gap", "totaltime", "dual_time", "lanczos_steps")},
          f"wall={time.perf_counter() - t0:.2f}")
