"""Full-size trajectory check of the multi-GPU path: the same C5 problem, R0 from the device generator, then the AL value
after 1, 10, 50, 200 and 400 inner iterations.  Run once on one GPU and once under torchrun; the printed JSON lines must
agree (R0 checksums exactly, AL values to ~1e-9 early on).
  python scripts/check_multigpu_solve.py [n] [edges]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/check_multigpu_solve.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from sdplrplus.jl_b200 import dist as spdist
from bench import SimpleData, generate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
edges = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * n
rank, world, local = spdist.init_process_group()
import torch
torch.cuda.set_device(local)
h = spdist.make_handle(sp.Handle)
asm, b, normC, E, gen_s = generate(sp, n, edges, 42)
data = SimpleData(n, n, b)
eng = sp.B200Engine(data, handle=h, asm=asm)
del asm
h.set_rank(10, 4)
h.fill_uniform(sp._lib.MAT_R, 12345)
R = eng.get_R()
out = {"world": world, "R_sum": float(R.sum()), "R_sq": float((R * R).sum()), "R_head": R[:2, :3].ravel().tolist(),
       "R_tail": R[-1, -3:].tolist(), "R_mid": R[n // 2 + 7, :2].tolist()}
del R
h.upload_vec(sp._lib.VEC_LAMBDA, np.zeros(n))
h.sigma = 2.0
eng.r = 10
fg = eng.fg()
out["fg"] = list(fg)
done = 0
for k in (1, 10, 50, 200, 400):
    last = eng.iterate(k - done)
    done = k
    out[f"it{k}"] = [last[0], last[1], last[4]]
import time
h.synchronize()
t0 = time.perf_counter()
d = eng.dual_obj(float(n), 100, None, 777)
h.synchronize()
out["dual"] = [d[0], d[1], d[2]]
out["dual_s"] = time.perf_counter() - t0   # SDPLRP_LANCZOS_DIST=1: row-partitioned Lanczos (must give the same dual value)
out["lanczos_dist"] = os.environ.get("SDPLRP_LANCZOS_DIST", "0")
if rank == 0:
    print(json.dumps(out), flush=True)
h.close()
