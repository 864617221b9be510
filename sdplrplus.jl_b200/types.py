"""Problem containers mirroring the reference's data model.

Reference: src/structs.jl:11-183 (SymLowRankMatrix, SDPData), LuxurySparse's
SparseMatrixCOO as used at src/preprocess.jl:4-16, and the classification
loop of SolverAuxiliary (src/structs.jl:296-332).

Python is 0-based; everything is converted to Julia's 1-based int64 only when
it crosses the C ABI (`assemble_sparse`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp


class SymLowRankMatrix:
    """A = B * Diagonal(D) * B' (src/structs.jl:11-24); B is n x s."""

    def __init__(self, D, B):
        self.D = np.ascontiguousarray(np.asarray(D, np.float64).reshape(-1))
        B = np.asarray(B, np.float64)
        if B.ndim == 1:
            B = B.reshape(-1, 1)
        self.B = np.asfortranarray(B)
        assert self.B.shape[1] == self.D.shape[0]

    @property
    def shape(self):
        n = self.B.shape[0]
        return (n, n)

    def toarray(self):
        return (self.B * self.D) @ self.B.T

    def norm(self, p=2):
        """norm(A, 2) (Frobenius) / norm(A, Inf) (max abs) -- src/structs.jl:61-82."""
        U = self.B * self.D
        if p == 2:
            # ||B D B'||_F^2 = tr((B'B D)(B'B D)) without forming the n x n matrix
            G = self.B.T @ self.B
            M = G * self.D  # G @ diag(D)
            return float(np.sqrt(max(np.trace(M @ M), 0.0)))
        if p == np.inf:
            res = 0.0
            for i in range(self.B.shape[0]):
                res = max(res, float(np.max(np.abs(U @ self.B[i, :]))))
            return res
        raise ValueError("undefined norm for Constraint")


class SparseMatrixCOO:
    """Coordinate-format symmetric matrix with insertion order preserved
    (LuxurySparse.SparseMatrixCOO as used by the reference; 0-based here).
    Duplicates are allowed and add up."""

    def __init__(self, rows, cols, vals, n):
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.n = int(n)

    @property
    def shape(self):
        return (self.n, self.n)

    def toarray(self):
        A = np.zeros((self.n, self.n))
        np.add.at(A, (self.rows, self.cols), self.vals)
        return A


class Diagonal:
    """Diagonal constraint (converted with sparse(A), src/structs.jl:307-309)."""

    def __init__(self, d):
        self.d = np.ascontiguousarray(d, dtype=np.float64)

    @property
    def shape(self):
        return (self.d.size, self.d.size)

    def toarray(self):
        return np.diag(self.d)


class ConstraintBatch:
    """Many COO constraint matrices stored back to back (vectorised stand-in
    for a Julia Vector of 10^7 one-entry SparseMatrixCOO objects).  Behaves like
    a sequence of SparseMatrixCOO for small problems/tests."""

    def __init__(self, offsets, rows, cols, vals, n):
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.rows = np.ascontiguousarray(rows, dtype=np.int64)
        self.cols = np.ascontiguousarray(cols, dtype=np.int64)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.n = int(n)

    def __len__(self):
        return self.offsets.size - 1

    def __getitem__(self, i):
        a, b = self.offsets[i], self.offsets[i + 1]
        return SparseMatrixCOO(self.rows[a:b], self.cols[a:b], self.vals[a:b], self.n)

    @staticmethod
    def diagonal_units(n, val=1.0):
        """e_i e_i' for i = 0..n-1 (the Diag(X) = 1 block of MaxCut etc.)."""
        idx = np.arange(n, dtype=np.int64)
        return ConstraintBatch(np.arange(n + 1, dtype=np.int64), idx, idx, np.full(n, val), n)


def _findnz_csc(A):
    """findnz of a SparseMatrixCSC: column-major, rows ascending."""
    A = sp.csc_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    n = A.shape[1]
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr))
    return A.indices.astype(np.int64), cols, A.data.astype(np.float64)


@dataclass
class SDPData:
    """src/structs.jl:150-183.  `As` may mix scipy sparse matrices (CSC),
    SparseMatrixCOO, Diagonal, SymLowRankMatrix and ConstraintBatch (which
    expands to len(batch) consecutive constraints)."""
    C: object
    As: Sequence
    b: np.ndarray
    constraint_types: Optional[np.ndarray] = None
    n: int = field(init=False)
    m: int = field(init=False)
    has_inequalities: bool = field(init=False)

    def __post_init__(self):
        self.n = int(self.C.shape[0])
        self.b = np.ascontiguousarray(self.b, dtype=np.float64)
        self.m = sum(len(A) if isinstance(A, ConstraintBatch) else 1 for A in self.As)
        assert self.b.size == self.m, (self.b.size, self.m)
        if self.constraint_types is None:
            self.constraint_types = np.zeros(self.m, dtype=bool)
        self.constraint_types = np.ascontiguousarray(self.constraint_types, dtype=bool)
        self.has_inequalities = bool(self.constraint_types.any())

    def matrices(self):
        """Flat list of the m constraint matrices (small problems only)."""
        out = []
        for A in self.As:
            if isinstance(A, ConstraintBatch):
                out.extend(A[i] for i in range(len(A)))
            else:
                out.append(A)
        return out


def b_vector(data):
    return data.b


def C_matrix(data):
    return data.C


def frobenius_norm(A):
    """norm(A, 2) in the reference = Frobenius norm."""
    if isinstance(A, SymLowRankMatrix):
        return A.norm(2)
    if isinstance(A, Diagonal):
        return float(np.linalg.norm(A.d))
    if isinstance(A, SparseMatrixCOO):
        return float(np.linalg.norm(sp.coo_matrix((A.vals, (A.rows, A.cols)), shape=A.shape).tocsc().data))
    if sp.issparse(A):
        return float(np.linalg.norm(sp.csc_matrix(A).data))
    return float(np.linalg.norm(np.asarray(A)))


@dataclass
class AssembledSparse:
    """What crosses the ABI into sdplrp_preprocess (all 1-based int64)."""
    n: int
    m: int
    mat_off: np.ndarray
    I: np.ndarray
    J: np.ndarray
    V: np.ndarray
    gids: np.ndarray
    lowrank: List  # (gid1, SymLowRankMatrix)


def assemble_sparse(data: SDPData) -> AssembledSparse:
    """The classification loop of SolverAuxiliary (src/structs.jl:303-332):
    sparse / diagonal A_i in order of appearance, then C if sparse, each as
    1-based triplets in findnz order; low-rank ones are listed separately."""
    Is, Js, Vs, lens, gids, lowrank = [], [], [], [], [], []
    gid = 0

    def push(rows, cols, vals, g):
        Is.append(rows); Js.append(cols); Vs.append(vals); lens.append(np.array([rows.size], np.int64)); gids.append(np.array([g], np.int64))

    def classify(A, g, what):
        if isinstance(A, SparseMatrixCOO):
            push(A.rows, A.cols, A.vals, g)
        elif isinstance(A, Diagonal):
            idx = np.arange(A.d.size, dtype=np.int64)  # sparse(Diagonal) stores all n diagonal entries
            push(idx, idx, A.d, g)
        elif isinstance(A, SymLowRankMatrix):
            lowrank.append((g, A))
        elif sp.issparse(A):
            r, c, v = _findnz_csc(A)
            push(r, c, v, g)
        else:
            raise TypeError(f"Currently only sparse/symmetric low-rank/diagonal {what} are supported.")

    for A in data.As:
        if isinstance(A, ConstraintBatch):
            k = len(A)
            Is.append(A.rows); Js.append(A.cols); Vs.append(A.vals)
            lens.append(np.diff(A.offsets)); gids.append(np.arange(gid + 1, gid + k + 1, dtype=np.int64))
            gid += k
        else:
            gid += 1
            classify(A, gid, "constraints")
    classify(data.C, data.m + 1, "objectives")
    if lens:
        lens_all = np.concatenate(lens)
        mat_off = np.concatenate([[0], np.cumsum(lens_all)]).astype(np.int64)
        I = np.concatenate(Is).astype(np.int64) + 1
        J = np.concatenate(Js).astype(np.int64) + 1
        V = np.concatenate(Vs).astype(np.float64)
        g = np.concatenate(gids).astype(np.int64)
    else:
        mat_off = np.zeros(1, np.int64); I = np.zeros(0, np.int64); J = np.zeros(0, np.int64)
        V = np.zeros(0, np.float64); g = np.zeros(0, np.int64)
    return AssembledSparse(data.n, data.m, mat_off, I, J, V, g, lowrank)


def structured_blocks(data: SDPData):
    """The sparse list of `assemble_sparse` as structured blocks for sdplrp_preprocess_blocks: a ConstraintBatch
    of one-entry diagonal matrices becomes ONE `DIAG` descriptor, a batch of {(i,j),(j,i)} pairs one `EDGES` descriptor, a CSC
    matrix goes over as its three CSC arrays, the identity as a bare descriptor; anything else falls back to a TRIPLETS block.
    Returns (blocks, lowrank) with lowrank = [(gid1, SymLowRankMatrix)]."""
    from . import _lib
    blocks, lowrank = [], []
    gid = 0

    def matrix_block(A, g, what):
        if isinstance(A, SparseMatrixCOO):
            blocks.append({"kind": _lib.BLOCK_TRIPLETS, "first_gid": g, "I": A.rows + 1, "J": A.cols + 1, "V": A.vals})
        elif isinstance(A, Diagonal):
            idx = np.arange(1, A.d.size + 1, dtype=np.int64)
            blocks.append({"kind": _lib.BLOCK_TRIPLETS, "first_gid": g, "I": idx, "J": idx, "V": A.d})
        elif isinstance(A, SymLowRankMatrix):
            lowrank.append((g, A))
        elif sp.issparse(A):
            M = sp.csc_matrix(A)
            M.sum_duplicates(); M.sort_indices()
            n = M.shape[0]
            if M.nnz == n and np.array_equal(M.indices, np.arange(n)) and np.all(M.data == 1.0) and np.array_equal(M.indptr, np.arange(n + 1)):
                blocks.append({"kind": _lib.BLOCK_IDENTITY, "first_gid": g})
            else:
                blocks.append({"kind": _lib.BLOCK_CSC, "first_gid": g, "I": M.indices.astype(np.int64) + 1,
                               "J": M.indptr.astype(np.int64) + 1, "V": M.data.astype(np.float64)})
        else:
            raise TypeError(f"Currently only sparse/symmetric low-rank/diagonal {what} are supported.")

    for A in data.As:
        if isinstance(A, ConstraintBatch):
            k = len(A)
            lens = np.diff(A.offsets)
            if k > 0 and np.all(lens == 1) and np.array_equal(A.rows, A.cols):
                pos = A.rows + 1
                blk = {"kind": _lib.BLOCK_DIAG, "first_gid": gid + 1, "count": k}
                if not np.array_equal(A.rows, np.arange(k)):
                    blk["I"] = pos
                if not np.all(A.vals == 1.0):
                    blk["V"] = A.vals
                blocks.append(blk)
            elif k > 0 and np.all(lens == 2) and np.array_equal(A.rows[0::2], A.cols[1::2]) and np.array_equal(A.cols[0::2], A.rows[1::2]) \
                    and np.array_equal(A.vals[0::2], A.vals[1::2]):
                blk = {"kind": _lib.BLOCK_EDGES, "first_gid": gid + 1, "count": k, "I": A.rows[0::2] + 1, "J": A.cols[0::2] + 1}
                if not np.all(A.vals == 1.0):
                    blk["V"] = A.vals[0::2].copy()
                blocks.append(blk)
            else:   # mixed batch: matrix by matrix
                for i in range(k):
                    matrix_block(A[i], gid + 1 + i, "constraints")
            gid += k
        else:
            gid += 1
            matrix_block(A, gid, "constraints")
    matrix_block(data.C, data.m + 1, "objectives")
    return blocks, lowrank
