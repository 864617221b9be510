"""Is the BDB' contraction of a SymLowRankMatrix (src/structs.jl:135-145) worth FP64 tensor cores?  Measurement for the waiver
in DESIGN.md: the projection X'B (n x r times n x s, the only dense contraction on the hot path) is timed at n = 10^7, r = 10
for s = 1, 4, 8 through the seam-level operator A(UU') on a problem whose only constraint is the low-rank matrix, and compared
with its algorithmic bytes.  Arithmetic intensity 2rs flops per 8(r + s) bytes of a row: 0.23 / 0.71 / 1.11 flop per byte --
against ~6 flop/byte where FP64 FMA throughput (40 TFLOP/s) would start to matter on a B200."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sps
import sdplrplus.jl_b200 as sp

n, r = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 10
rng = np.random.default_rng(0)
out = []
for s in (1, 4, 8):
    C = sps.identity(n, format="csc")
    B = rng.standard_normal((n, s))
    data = sp.SDPData(C, [sp.SymLowRankMatrix(np.ones(s), B)], np.zeros(1))
    h = sp.Handle(device=0)
    eng = sp.B200Engine(data, handle=h)
    eng.init_vars(r, 2.0 * rng.random((n, r)) - 1.0, np.zeros(1), 2.0, 4)
    for _ in range(3):
        h.A_uu()
    import time
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        h.A_uu()     # returns two host doubles: synchronous
    ms = 1e3 * (time.perf_counter() - t0) / K
    nbytes = 8.0 * n * (r + s) + 8.0 * n * r   # X and B once for the projection; the identity objective streams X once more
    out.append({"s": s, "ms": ms, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / ms / 1e6, "flops_per_byte": 2.0 * r * s / (8.0 * (r + s)),
                "timing": "wall clock around 10 synchronous A(UU') calls"})
    h.close()
print(json.dumps(out))
