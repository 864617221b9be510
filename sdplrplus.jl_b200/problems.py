"""SDP problem constructors and synthetic graph generators for the BASELINE configs.

Reference: test/problem.jl:16-236 (== exps/problems.jl): maxcut, lovasz_theta,
minimum_bisection, cutnorm, mu_conductance, mu_conductance_ineq.  Each returns
(C, As, bs[, constraint_types]) ready for SDPData / sdplr(); the ORDER of `As`
is the reference's because it fixes matptr / nzind / lambda.

Everything is vectorised numpy (no per-constraint Python objects) so the
10M-vertex config can be assembled on the host in seconds.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .types import ConstraintBatch, SparseMatrixCOO, SymLowRankMatrix


def _check_undirected(A):
    A = sp.csc_matrix(A)
    if A.shape[0] <= 5000:
        assert (abs(A - A.T)).nnz == 0, "Only undirected graphs supported now."
    return A


def maxcut(A):
    """minimize -1/4 <L, X>  s.t. Diag(X) = 1   (test/problem.jl:16-31)"""
    A = _check_undirected(A)
    n = A.shape[0]
    d = np.asarray(A.sum(axis=1)).reshape(-1)
    L = sp.csc_matrix(sp.diags(d) - A)
    L = L * -0.25
    As = [ConstraintBatch.diagonal_units(n)]
    bs = np.ones(n)
    return sp.csc_matrix(L), As, bs


def lovasz_theta(A):
    """minimize -<11', X> s.t. Tr X = 1, X_ij = 0 on edges   (test/problem.jl:43-65)"""
    A = _check_undirected(A)
    n = A.shape[0]
    C = SymLowRankMatrix(-np.ones(1), np.ones((n, 1)))
    A.sort_indices()
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
    rows = A.indices.astype(np.int64)
    up = rows < cols   # findnz order (column-major), one constraint per stored entry with i < j
    dg = rows == cols
    # constraints in findnz order: i<j -> COO {(i,j),(j,i)}; i==j -> COO {(i,i)}
    keep = up | dg
    r, c = rows[keep], cols[keep]
    isdiag = r == c
    cnt = np.where(isdiag, 1, 2).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(cnt)])
    tot = int(offsets[-1])
    er = np.empty(tot, np.int64); ec = np.empty(tot, np.int64)
    first = offsets[:-1]
    er[first] = r; ec[first] = c
    second = first[~isdiag] + 1
    er[second] = c[~isdiag]; ec[second] = r[~isdiag]
    As = []
    if r.size:
        As.append(ConstraintBatch(offsets, er, ec, np.ones(tot), n))
    As.append(sp.identity(n, format="csc"))
    bs = np.concatenate([np.zeros(r.size), [1.0]])
    return C, As, bs


def minimum_bisection(A):
    """minimize 1/4 <L, X> s.t. Diag(X) = 1, 1'X1 = 0   (test/problem.jl:78-95)"""
    A = _check_undirected(A)
    n = A.shape[0]
    d = np.asarray(A.sum(axis=1)).reshape(-1)
    L = sp.csc_matrix(sp.diags(d) - A) / 4.0
    As = [ConstraintBatch.diagonal_units(n), SymLowRankMatrix(np.ones(1), np.ones((n, 1)))]
    bs = np.concatenate([np.ones(n), [0.0]])
    return sp.csc_matrix(L), As, bs


def cutnorm(A):
    """C = -1/2 [0 A; A' 0], Diag(X) = 1   (test/problem.jl:97-113)"""
    A = sp.csc_matrix(A)
    m, n = A.shape
    B = sp.bmat([[sp.csc_matrix((m, m)), A], [A.T, sp.csc_matrix((n, n))]], format="csc") / 2.0
    N = m + n
    As = [ConstraintBatch.diagonal_units(N)]
    bs = np.ones(N)
    return sp.csc_matrix(-B), As, bs


def mu_conductance_ub(volG, mu):
    return (1 - mu) / (mu * volG)


def mu_conductance_lb(volG, mu):
    return mu / ((1 - mu) * volG)


def mu_conductance(A, mu):
    """3n-lifted mu-conductance SDP   (test/problem.jl:139-179)"""
    A = _check_undirected(A)
    n = A.shape[0]
    d = np.asarray(A.sum(axis=1)).reshape(-1)
    volG = d.sum()
    L = sp.csc_matrix(sp.diags(d) - A)
    N = 3 * n
    padded_d = np.concatenate([d, np.zeros(2 * n)])
    Dcoo = sp.csc_matrix(sp.diags(d)).tocoo()     # findnz(D): stored (structurally non-zero) diagonal entries
    Lcoo = L.tocoo()
    padded_D = sp.csc_matrix((Dcoo.data, (Dcoo.row, Dcoo.col)), shape=(N, N))
    padded_L = sp.csc_matrix((Lcoo.data, (Lcoo.row, Lcoo.col)), shape=(N, N))
    ub, lb = mu_conductance_ub(volG, mu), mu_conductance_lb(volG, mu)
    idx = np.arange(n, dtype=np.int64)
    off = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
    rows_ub = np.stack([idx, idx + n], axis=1).reshape(-1)
    rows_lb = np.stack([idx, idx + 2 * n], axis=1).reshape(-1)
    As = [padded_D,
          SymLowRankMatrix(np.ones(1), padded_d.reshape(-1, 1)),
          ConstraintBatch(off, rows_ub, rows_ub, np.ones(2 * n), N),
          ConstraintBatch(off, rows_lb, rows_lb, np.tile([1.0, -1.0], n), N)]
    bs = np.concatenate([[1.0, 0.0], np.full(n, ub), np.full(n, lb)])
    return padded_L, As, bs


def mu_conductance_ineq(A, mu):
    """n x n mu-conductance SDP with native inequality constraints (test/problem.jl:196-236)"""
    A = _check_undirected(A)
    n = A.shape[0]
    d = np.asarray(A.sum(axis=1)).reshape(-1)
    D = sp.csc_matrix(sp.diags(d))
    volG = d.sum()
    L = sp.csc_matrix(sp.diags(d) - A)
    ub, lb = mu_conductance_ub(volG, mu), mu_conductance_lb(volG, mu)
    As = [D, SymLowRankMatrix(np.ones(1), d.reshape(-1, 1)),
          ConstraintBatch.diagonal_units(n, 1.0), ConstraintBatch.diagonal_units(n, -1.0)]
    bs = np.concatenate([[1.0, 0.0], np.full(n, ub), np.full(n, -lb)])
    types = np.concatenate([[False, False], np.ones(2 * n, dtype=bool)])
    return L, As, bs, types


# ---------------------------------------------------------------------------
# synthetic graphs for the BASELINE configs
# ---------------------------------------------------------------------------
def make_random_graph(n, p, rng):
    """test/runtests.jl:30-36 (edge iff symmetrised uniform > p), NumPy RNG."""
    X = rng.random((n, n))
    X = (X + X.T) / 2
    A = (X > p).astype(np.float64)
    np.fill_diagonal(A, 0.0)
    return sp.csc_matrix(A)


def _sym_from_edges(n, u, v, w=None):
    keep = u != v
    u, v = u[keep], v[keep]
    if w is not None:
        w = w[keep]
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    key = lo.astype(np.int64) * n + hi
    key, first = np.unique(key, return_index=True)
    lo, hi = key // n, key % n
    vals = np.ones(lo.size) if w is None else w[first]
    A = sp.coo_matrix((np.concatenate([vals, vals]), (np.concatenate([lo, hi]), np.concatenate([hi, lo]))), shape=(n, n))
    return sp.csc_matrix(A)


def gnm_graph(n, M, seed):
    """Uniform simple graph with ~M edges (C1: n=800, M=19176, the G1 shape)."""
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n, size=int(M * 1.15) + 16)
    v = rng.integers(0, n, size=u.size)
    A = _sym_from_edges(n, u, v)
    if A.nnz // 2 > M:  # trim to exactly M edges, deterministically
        T = sp.triu(A, k=1).tocoo()
        sel = np.sort(rng.permutation(T.nnz)[:M])
        A = _sym_from_edges(n, T.row[sel].astype(np.int64), T.col[sel].astype(np.int64))
    return A


def erdos_renyi(n, p, seed):
    rng = np.random.default_rng(seed)
    M = rng.binomial(n * (n - 1) // 2, p)
    return gnm_graph(n, int(M), seed + 1)


def powerlaw_graph(n, target_edges, seed, exponent=2.3):
    """Chung-Lu style power-law graph (C5): expected degree w_i ~ (i+i0)^(-1/(exponent-1)),
    endpoints drawn with probability proportional to w, symmetrised, de-duplicated, loop-free."""
    rng = np.random.default_rng(seed)
    gamma = 1.0 / (exponent - 1.0)
    i0 = max(1.0, n * 1e-6 * 10)
    w = (np.arange(n, dtype=np.float64) + i0) ** (-gamma)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    M = int(target_edges * 1.03)
    u = np.searchsorted(cdf, rng.random(M)).astype(np.int64)
    v = np.searchsorted(cdf, rng.random(M)).astype(np.int64)
    perm = rng.permutation(n)  # hide the degree ordering in the vertex labels
    return _sym_from_edges(n, perm[np.minimum(u, n - 1)], perm[np.minimum(v, n - 1)])


def powerlaw_maxcut_assembled(n, target_edges, seed, device=None, exponent=2.3, keep_on_device=False):
    """MaxCut on a Chung-Lu power-law graph, assembled straight into the ABI's triplet form
    (n one-entry diagonal constraints, then C = -1/4 L in CSC `findnz` order) without scipy:
    torch does the sampling / sort / unique, on the GPU when `device` is a CUDA device
    (the 10M-vertex BASELINE config takes ~170 s through scipy and a few seconds this way).
    Returns (AssembledSparse, b, normC)."""
    import torch
    from .types import AssembledSparse
    dev = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    gamma = 1.0 / (exponent - 1.0)
    i0 = max(1.0, n * 1e-5)
    # the cumulative weights are summed on the host: a CUDA cumsum is not bitwise reproducible from run to run, and a cdf
    # that moves in its last bits moves a handful of the 8e7 sampled endpoints (seen as 1e-9 relative differences in <C, RR'>)
    w = (np.arange(n, dtype=np.float64) + i0) ** (-gamma)
    cdf_h = np.cumsum(w)
    cdf_h /= cdf_h[-1]
    cdf = torch.from_numpy(cdf_h).to(dev)
    del w, cdf_h
    M = int(target_edges * 1.03)
    u = torch.searchsorted(cdf, torch.rand(M, dtype=torch.float64, device=dev, generator=g)).clamp_(max=n - 1)
    v = torch.searchsorted(cdf, torch.rand(M, dtype=torch.float64, device=dev, generator=g)).clamp_(max=n - 1)
    perm = torch.randperm(n, device=dev, generator=g)
    u, v = perm[u], perm[v]
    del cdf, perm
    keep = u != v
    lo, hi = torch.minimum(u, v)[keep], torch.maximum(u, v)[keep]
    del u, v, keep
    key = torch.unique(lo * n + hi)          # sorted, de-duplicated undirected edges
    del lo, hi
    lo, hi = key // n, key % n
    E = int(key.numel())
    deg = torch.bincount(lo, minlength=n) + torch.bincount(hi, minlength=n)
    diag = torch.arange(n, dtype=torch.int64, device=dev)
    # column-major keys (col * n + row) of C: both triangles plus the full diagonal
    ck = torch.cat([hi * n + lo, lo * n + hi, diag * n + diag])
    del key, lo, hi
    ck, _ = torch.sort(ck)
    col, row = ck // n, ck % n
    del ck
    V = torch.where(row == col, -0.25 * deg[row].to(torch.float64), torch.full((1,), 0.25, dtype=torch.float64, device=dev))
    normC = float(torch.sqrt(torch.sum(V * V)).item())
    I_t = torch.cat([diag + 1, row + 1])
    J_t = torch.cat([diag + 1, col + 1])
    V_t = torch.cat([torch.ones(n, dtype=torch.float64, device=dev), V])
    nnzC = int(V.numel())
    mat_off = np.concatenate([np.arange(n + 1, dtype=np.int64), [n + nnzC]]).astype(np.int64)
    gids = np.arange(1, n + 2, dtype=np.int64)
    if keep_on_device and dev.type == "cuda":
        # direct device construction: the triplets stay where they were built and go to
        # sdplrp_preprocess_device; the host arrays of the ABI's other form are not materialised
        empty_i, empty_f = np.empty(0, np.int64), np.empty(0, np.float64)
        asm = AssembledSparse(n, n, mat_off, empty_i, empty_i, empty_f, gids, [])
        asm.device_triplets = (I_t.contiguous(), J_t.contiguous(), V_t.contiguous())
        return asm, np.ones(n), normC, E
    asm = AssembledSparse(n, n, mat_off, I_t.cpu().numpy(), J_t.cpu().numpy(), V_t.cpu().numpy(), gids, [])
    return asm, np.ones(n), normC, E
