"""How much of the gather pass of C5 could hit L2?  Degree statistics of the actual benchmark graph family
(problems.powerlaw_maxcut_assembled's sampler, restated with numpy so that it runs without a GPU): for a hub prefix of
k vertices (hub-first order) the share of gathers that target it, and the split of the nonzeros into
(hub row, hub col) / (hub, tail) + (tail, hub) / (tail, tail).  Feeds DESIGN.md section 8.
  python scripts/hub_tail_analysis.py [n] [edges]"""
import sys
import numpy as np

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
edges = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * n
rng = np.random.default_rng(42)
gamma = 1.0 / (2.3 - 1.0)
i0 = max(1.0, n * 1e-5)
w = (np.arange(n, dtype=np.float64) + i0) ** (-gamma)
cdf = np.cumsum(w); cdf /= cdf[-1]
M = int(edges * 1.03)
u = np.minimum(np.searchsorted(cdf, rng.random(M)), n - 1)
v = np.minimum(np.searchsorted(cdf, rng.random(M)), n - 1)
keep = u != v
lo, hi = np.minimum(u, v)[keep], np.maximum(u, v)[keep]
key = np.unique(lo.astype(np.int64) * n + hi)
lo, hi = key // n, key % n
E = key.size
deg = np.bincount(lo, minlength=n) + np.bincount(hi, minlength=n)
order = np.argsort(-deg, kind="stable")
rank_of = np.empty(n, np.int64); rank_of[order] = np.arange(n)
rl, rh = rank_of[lo], rank_of[hi]              # hub-first labels of both endpoints
sdeg = deg[order]
cum = np.cumsum(sdeg) / (2.0 * E)
print(f"n={n} E={E} mean degree={2*E/n:.2f} max degree={sdeg[0]} median={int(np.median(deg))}")
print("L2 turnover argument: rows gathered >= 170 times per pass:", int((sdeg >= 170).sum()), "vertices holding",
      f"{cum[max((sdeg >= 170).sum() - 1, 0)]:.3f} of the gathers")
print(f"{'hub rows':>10} {'MB (80 B rows)':>15} {'gathers into hub':>17} {'(H,H)':>7} {'(H,T)+(T,H)':>12} {'(T,T)':>7}")
for k in (100_000, 200_000, 400_000, 800_000, 1_200_000, 1_600_000):
    if k >= n:
        break
    a, b = rl < k, rh < k
    hh = np.count_nonzero(a & b) / E
    tt = np.count_nonzero(~a & ~b) / E
    print(f"{k:>10} {k*80/1e6:>15.0f} {cum[k-1]:>17.3f} {hh:>7.3f} {1-hh-tt:>12.3f} {tt:>7.3f}")
