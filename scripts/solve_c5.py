"""Time-to-tolerance of the full solve on the C5 workload (driver = sdplrplus.jl_b200.solver._sdplr, reference defaults)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from bench import SimpleData, generate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
edges = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * n
asm, b, normC, E, gen_s = generate(sp, n, edges, 42)
data = SimpleData(n, n, b)
data.C = None
h = sp.Handle(device=0)
eng = sp.B200Engine(data, handle=h, asm=asm)
cfg = sp.BurerMonteiroConfig(prior_trace_bound=float(n), printlevel=1, printfreq=5.0, maxtime=600.0, dataset=f"C5 n={n}", seed=0,
                             lanczos_host_rng=False)
stats = sp.solver.SolverStats()
import sdplrplus.jl_b200.solver as S
S.frobenius_norm = lambda C: normC   # ||C||_F of the assembled objective (the triplets never exist as a scipy matrix here)
res = S._sdplr(data, eng, cfg, stats, 10, np.random.default_rng(0))
print(json.dumps({k: res[k] for k in ("iter", "majoriter", "obj", "primal_vio", "max_dual_value", "min_duality_gap", "totaltime", "dual_time",
                                       "primaltime", "lanczos_steps", "r")}))
