"""Debug helper: compares the seam-level products of the two SpMM kernels on a tiny MaxCut."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdplrplus.jl_b200 as sp
from sdplrplus.jl_b200 import _lib

P = sp.problems
n, r = int(sys.argv[1]) if len(sys.argv) > 1 else 5, int(sys.argv[2]) if len(sys.argv) > 2 else 2
C, As, bs = P.maxcut(P.erdos_renyi(n, 0.5, 1))
data = sp.SDPData(C, As, bs)
Rt0 = 2 * np.random.default_rng(0).random((n, r)) - 1
Cd = C.toarray()
for kern in (0, 1):
    for rel in (0, 1):
        h = sp.Handle(device=0)
        h.set_option("spmm_kernel", kern); h.set_option("relabel", rel)
        ge = sp.B200Engine(data, handle=h)
        ge.init_vars(r, Rt0, np.zeros(n), 2.0, 4)
        L, obj = ge.f()
        raw = ge.get_pvio_raw()
        y = np.concatenate([np.arange(1, n + 1) * 0.1, [1.0]])
        h._check(h.lib.sdplrp_At_preprocess(h._h, _lib._f64(y)[1]))
        h.upload_mat(_lib.MAT_W0, Rt0)
        h._check(h.lib.sdplrp_At_left(h._h, _lib.MAT_W0, _lib.MAT_W1))
        Y = h.download_mat(_lib.MAT_W1)
        S = Cd + np.diag(y[:n])
        print(f"kernel={kern} relabel={rel} obj={obj:.6f} expect={np.sum((Cd @ Rt0) * Rt0):.6f}  At_left err={np.abs(Y - S @ Rt0).max():.3e}")
        gn2, pn2 = ge.g()
        G = ge.get_G()
        yv = ge.get_y()
        Sg = Cd + np.diag(yv[:n])
        print(f"    g err={np.abs(G - 2 * Sg @ Rt0).max():.3e}")
        h.close()
