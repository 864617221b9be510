"""Diagnostic for the full-size multi-GPU path: are the per-rank problem copies identical, and which constraint slots of
primal_vio_raw differ from the exact values after f!?  (run under torchrun, 2 ranks)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from sdplrplus.jl_b200 import dist as spdist
from bench import SimpleData, generate

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
edges = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * n
rank, world, local = spdist.init_process_group()
import torch, torch.distributed as tdist
torch.cuda.set_device(local)
h = spdist.make_handle(sp.Handle)
asm, b, normC, E, gen_s = generate(sp, n, edges, 42)
I64 = asm.I.astype(np.uint64); J64 = asm.J.astype(np.uint64)
with np.errstate(over="ignore"):
    chk = int(((I64 * np.uint64(0x9E3779B97F4A7C15)) ^ (J64 * np.uint64(0xD6E8FEB86659FD93))).sum(dtype=np.uint64))
    chk_v = float(asm.V.sum())
info = torch.tensor([float(E), float(chk % (1 << 52)), chk_v, float(asm.I.size)], dtype=torch.float64, device="cuda")
allinfo = [torch.zeros_like(info) for _ in range(world)]
tdist.all_gather(allinfo, info)
data = SimpleData(n, n, b)
eng = sp.B200Engine(data, handle=h, asm=asm)
del asm, I64, J64
h.set_rank(10, 4)
h.fill_uniform(sp._lib.MAT_R, 12345)
R = eng.get_R()
truth = (R * R).sum(1) - 1.0
del R
h.upload_vec(sp._lib.VEC_LAMBDA, np.zeros(n))
h.sigma = 2.0
eng.r = 10
fg = eng.fg()
raw = eng.get_pvio_raw()
bdev = h.download_vec(sp._lib.VEC_B, n)
bad = np.nonzero(np.abs(raw[:n] - truth) > 1e-9)[0]
out = {"rank": rank, "row_range": h.row_range(), "graphs": [t.tolist() for t in allinfo], "fg": list(fg), "n_bad": int(bad.size),
       "bad_head": bad[:10].tolist(), "raw_bad": raw[bad[:10]].tolist(), "truth_bad": truth[bad[:10]].tolist(),
       "b_not_one": int((bdev != 1.0).sum()), "pn2_from_raw": float((raw[:n] ** 2).sum()), "pn2_truth": float((truth ** 2).sum()),
       "obj_slot": float(raw[n])}
print(json.dumps(out), flush=True)
h.close()
