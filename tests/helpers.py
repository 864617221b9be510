"""Shared helpers of the parity suite (design follows the reference's test/coreop.jl)."""
import json
import os

import numpy as np
import scipy.sparse as sps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAP_KEYS = ["triu_colptr", "triu_rowval", "matptr", "nzind", "nzval_one", "nzval_two", "full_colptr", "full_rowval", "mapped"]


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def g1_graph():
    z = np.load(os.path.join(GOLDEN, "g1_graph.npz"))
    n = int(z["n"])
    indptr = z["indptr"].astype(np.int64)
    A = sps.csc_matrix((np.ones(indptr[-1]), z["indices"].astype(np.int64), indptr), shape=(n, n))
    return A


def k2_graph():
    return sps.csc_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))


def p3_graph():
    return sps.csc_matrix(np.array([[0.0, 1, 0], [1, 0, 1], [0, 1, 0]]))


def dense_of(A):
    return A.toarray() if hasattr(A, "toarray") else np.asarray(A)


def dense_primal_vio(data, Rt):
    """test/coreop.jl:8-16 -- Rt is (n, r): [<A_i, RR'> - b_i ; <C, RR'>]."""
    X = Rt @ Rt.T
    As = data.matrices()
    out = np.zeros(data.m + 1)
    for i, A in enumerate(As):
        out[i] = np.sum(dense_of(A) * X) - data.b[i]
    out[data.m] = np.sum(dense_of(data.C) * X)
    return out


def dense_S(data, y):
    """test/coreop.jl:122-127"""
    S = y[data.m] * dense_of(data.C)
    for i, A in enumerate(data.matrices()):
        S = S + y[i] * dense_of(A)
    return S


def families(sp):
    P = sp.problems
    return [
        ("maxcut", P.maxcut), ("lovasz_theta", P.lovasz_theta), ("minimum_bisection", P.minimum_bisection),
        ("cutnorm", P.cutnorm), ("mu_conductance_0.01", lambda A: P.mu_conductance(A, 0.01)),
        ("mu_conductance_0.05", lambda A: P.mu_conductance(A, 0.05)), ("mu_conductance_0.1", lambda A: P.mu_conductance(A, 0.1)),
    ]


# test/coreop.jl:46-47: enumerate(Iterators.product([5,8,12],[0.4,0.7],[2,3])), first index fastest
COMBOS = [(seed + 1, n, p, r) for seed, (r, p, n) in
          enumerate((r, p, n) for r in (2, 3) for p in (0.4, 0.7) for n in (5, 8, 12))]


def make_case(sp, fam_fn, seed, n, p, r, ineq=False):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(50):  # the reference's generator can return an empty graph for tiny n; redraw deterministically
        A = sp.problems.make_random_graph(n, p, rng)
        if A.nnz > 0:
            break
    out = fam_fn(A)
    if len(out) == 4:
        C, As, bs, types = out
    else:
        (C, As, bs), types = out, None
    data = sp.SDPData(C, As, bs, types)
    Rt0 = 2.0 * rng.random((data.n, r)) - 1.0
    return data, Rt0, rng
