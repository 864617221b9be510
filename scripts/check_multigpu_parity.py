"""Multi-GPU parity: run under torchrun with N ranks (one GPU each).  Every rank drives its own handle through the same
call sequence (SPMD); rank 0 also runs the CPU oracle and compares the scalars every call returns and the downloaded
state, for problem families that exercise per-row constraints, general sparse constraints and low-rank terms."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from sdplrplus.jl_b200 import dist as spdist
from oracle import pyoracle
import torch

rank, world, local = spdist.init_process_group()
torch.cuda.set_device(local)
P = sp.problems
import scipy.sparse as sps
cases = {
    "maxcut_powerlaw": P.maxcut(P.powerlaw_graph(3000, 24000, 5)),
    "maxcut_g1shape": P.maxcut(P.gnm_graph(800, 19176, 1)),
    "bisection": P.minimum_bisection(P.erdos_renyi(600, 0.02, 3)),
    "lovasz": P.lovasz_theta(P.erdos_renyi(150, 0.06, 2)),
    "cutnorm": P.cutnorm(sps.random(120, 90, density=0.08, random_state=4, data_rvs=np.random.default_rng(4).standard_normal, format="csc")),
}
fail = 0
for name, (C, As, bs) in cases.items():
    for relabel in (1, 0):
        data = sp.SDPData(C, As, bs)
        r = 6
        Rt0 = 2 * np.random.default_rng(0).random((data.n, r)) - 1
        h = spdist.make_handle(sp.Handle)
        h.set_option("relabel", relabel)
        ge = sp.B200Engine(data, handle=h)
        ge.init_vars(r, Rt0, np.zeros(data.m), 2.0, 4)
        oe = pyoracle.OracleEngine(data); oe.init_vars(r, Rt0, np.zeros(data.m), 2.0, 4)
        def close(a, b, tol, what):
            global fail
            a, b = np.asarray(a, float), np.asarray(b, float)
            err = float(np.max(np.abs(a - b))) if a.size else 0.0
            if not err <= tol * max(1.0, float(np.max(np.abs(b))) if b.size else 1.0):
                fail += 1
                if rank == 0: print(f"  MISMATCH {name} relabel={relabel} {what}: err {err:.3e}")
        close(ge.fg(), oe.fg(), 1e-10, "fg")
        close(ge.get_G(), oe.get_G(), 1e-10, "G")
        for it in range(4):
            dg, do = ge.lbfgs_dir(), oe.lbfgs_dir()
            if math.isnan(do) or do >= 0:
                ge.use_gradient_direction(); oe.use_gradient_direction()
            else:
                close(dg, do, 1e-8, f"descent it={it}")
            bqg, bqo = ge.linesearch_coeffs(), oe.linesearch_coeffs()
            close(bqg, bqo, 1e-8, f"bq it={it}")
            ao, _ = sp.pick_alpha(bqo, 1.0)
            objg, gn2, pn2 = ge.step_g(ao)
            objo = oe.step(ao); ogn2, opn2 = oe.g()
            close([objg, math.sqrt(gn2), math.sqrt(pn2)], [objo, math.sqrt(ogn2), math.sqrt(opn2)], 1e-7, f"step_g it={it}")
            ge.lbfgs_update(ao); oe.lbfgs_update(ao)
        close(ge.get_R(), oe.get_R(), 1e-7, "R")
        close(ge.get_pvio_raw(), oe.get_pvio_raw(), 1e-7, "raw")
        close(ge.get_lambda(), oe.get_lambda(), 1e-9, "lambda")
        v0 = np.random.default_rng(1).standard_normal(data.n)
        dgv, eg, _ = ge.dual_obj(float(data.n), 100, v0)
        dov, eo, _ = oe.dual_obj(float(data.n), 100, v0)
        close([eg], [eo], 1e-4, "mineig")
        close([dgv], [dov], 1e-4, "dual")
        ge.dual_update(); oe.dual_update()
        close(ge.get_lambda(), oe.get_lambda(), 1e-7, "lambda after dual update")
        if rank == 0: print(f"{name:18s} relabel={relabel} world={world} rows={h.row_range()} ok so far (fail={fail})", flush=True)
        h.close()
spdist.barrier()
if rank == 0: print("MULTIGPU PARITY", "FAILED" if fail else "OK", flush=True)
sys.exit(1 if fail else 0)
