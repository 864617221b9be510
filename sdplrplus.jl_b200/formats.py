"""On-disk formats either side of the hot path.

Reference: exps/data_utils.jl:22-152 (SDPA and SDPLR-1.03 problem writers, the SDPLR initial-solution
file), exps/data_utils.jl:212-243 (`read_gset`), exps/data_preprocess.jl:85-116 / exps/data_utils.jl:1-13
(`read_graph`: MATLAB v7.3 `.mat` files holding one sparse adjacency matrix).

The writers emit what the reference's writers emit, entry for entry (`findnz(triu(A))` order: column-major
for CSC inputs, stored order for COO inputs; numbers printed the way Julia prints a Float64).  Readers for the
same formats are provided so that files round-trip and so that problems prepared for SDPLR-1.03 / CSDP can be
fed to this solver.  Everything here is host-side Python: none of it is on the per-iteration path.
"""
from __future__ import annotations

import decimal
import os
import struct

import numpy as np
import scipy.sparse as sps

from .types import ConstraintBatch, Diagonal, SparseMatrixCOO, SymLowRankMatrix, _findnz_csc


# ---------------------------------------------------------------------------------------------------------
# number formatting: Julia's `print(::Float64)` (shortest round-trip digits; fixed notation for 1e-4 <= |x| < 1e6)
# ---------------------------------------------------------------------------------------------------------
def jl_float(x):
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign, digits, exp = decimal.Decimal(repr(x)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:
        digits.pop(); exp += 1
    nd = len(digits)
    e10 = nd + exp - 1                      # x = d.ddd * 10^e10
    ds = "".join(map(str, digits))
    if -4 <= e10 < 6:
        if exp >= 0:
            body = ds + "0" * exp + ".0"
        elif -exp < nd:
            body = ds[: nd + exp] + "." + ds[nd + exp:]
        else:
            body = "0." + "0" * (-exp - nd) + ds
    else:
        body = ds[0] + "." + (ds[1:] if nd > 1 else "0") + "e" + str(e10)
    return ("-" if sign else "") + body


# ---------------------------------------------------------------------------------------------------------
# triu(findnz) of the matrix kinds the solver accepts (1-based rows/cols)
# ---------------------------------------------------------------------------------------------------------
def _triu_entries(A):
    """(rows, cols, vals), 1-based, in the order of `findnz(triu(A))`: CSC column-major (exps/data_utils.jl:41,46);
    COO in stored order (src/preprocess.jl:4-16)."""
    if isinstance(A, SparseMatrixCOO):
        keep = A.rows <= A.cols
        return A.rows[keep] + 1, A.cols[keep] + 1, A.vals[keep]
    if isinstance(A, Diagonal):
        idx = np.arange(1, A.d.size + 1, dtype=np.int64)
        return idx, idx, A.d
    if sps.issparse(A):
        r, c, v = _findnz_csc(sps.triu(sps.csc_matrix(A), format="csc"))
        return r + 1, c + 1, v
    raise TypeError("Only sparse and low-rank matrices are supported in SDPLR.")  # exps/data_utils.jl:82-86


def _flatten(As):
    out = []
    for A in As:
        if isinstance(A, ConstraintBatch):
            out.extend(A[i] for i in range(len(A)))
        else:
            out.append(A)
    return out


def _header(f, n, m, bs):
    f.write(f"{m}\n")      # number of constraint matrices
    f.write("1\n")         # number of blocks in the SDP
    f.write(f"{n}\n")      # sizes of the blocks
    f.write("".join(jl_float(b) + " " for b in bs) + "\n")


# ---------------------------------------------------------------------------------------------------------
# SDPA sparse format (exps/data_utils.jl:22-51)
# ---------------------------------------------------------------------------------------------------------
def write_problem_sdpa(path, C, As, bs):
    """write_problem_sdpa: `0 1 i j -C_ij` for triu(C), then `k 1 i j v` for triu(A_k).  Sparse matrices only."""
    As = _flatten(As)
    n = C.shape[0]
    with open(path, "w") as f:
        _header(f, n, len(As), bs)
        for i, j, v in zip(*_triu_entries(C)):
            f.write(f"0 1 {i} {j} {jl_float(-v)}\n")
        for k, A in enumerate(As, 1):
            for i, j, v in zip(*_triu_entries(A)):
                f.write(f"{k} 1 {i} {j} {jl_float(v)}\n")


def read_problem_sdpa(path):
    """Single-block SDPA sparse file -> (C, As, bs) with the sign convention of write_problem_sdpa (C = -F0);
    matrices come back as symmetric CSC."""
    with open(path) as f:
        lines = [ln.split("*")[0].split('"')[0].strip() for ln in f]
    lines = [ln for ln in lines if ln]
    m = int(lines[0].split()[0])
    nblocks = int(lines[1].split()[0])
    sizes = [int(t) for t in lines[2].replace(",", " ").replace("(", " ").replace(")", " ").replace("{", " ").replace("}", " ").split()]
    if nblocks != 1 or len(sizes) != 1:
        raise ValueError("only single-block SDPA problems are supported")
    n = abs(sizes[0])
    bs = np.array([float(t) for t in lines[3].replace(",", " ").replace("{", " ").replace("}", " ").split()], dtype=np.float64)
    if bs.size != m:
        raise ValueError("SDPA: right-hand side has the wrong length")
    ent = np.array([[float(t) for t in ln.split()] for ln in lines[4:]], dtype=np.float64).reshape(-1, 5)
    mats = []
    for k in range(m + 1):
        e = ent[ent[:, 0] == k]
        i, j, v = e[:, 2].astype(np.int64) - 1, e[:, 3].astype(np.int64) - 1, e[:, 4]
        off = i != j
        A = sps.coo_matrix((np.r_[v, v[off]], (np.r_[i, j[off]], np.r_[j, i[off]])), shape=(n, n)).tocsc()
        mats.append(A)
    return -mats[0], mats[1:], bs


# ---------------------------------------------------------------------------------------------------------
# SDPLR-1.03 format (exps/data_utils.jl:53-124)
# ---------------------------------------------------------------------------------------------------------
def _write_matrix_sdplr(f, A, mid):
    if isinstance(A, SymLowRankMatrix):
        f.write(f"{mid} 1 l {A.B.shape[1]}\n")       # matrix id, block id, low rank, rank
        for d in A.D:
            f.write(jl_float(d) + "\n")
        for j in range(A.B.shape[1]):                # B in column-major order
            for i in range(A.B.shape[0]):
                f.write(jl_float(A.B[i, j]) + "\n")
        return
    r, c, v = _triu_entries(A)
    f.write(f"{mid} 1 s {r.size}\n")
    for i, j, x in zip(r, c, v):
        f.write(f"{i} {j} {jl_float(x)}\n")


def write_problem_sdplr(path, C, As, bs):
    """write_problem_sdplr: header, the ignored `1` line, then C (id 0) and every A_i as sparse (`s`) or
    low-rank (`l`) sections."""
    As = _flatten(As)
    n = C.shape[0]
    with open(path, "w") as f:
        _header(f, n, len(As), bs)
        f.write("1\n")                                # this line is currently ignored
        _write_matrix_sdplr(f, C, 0)
        for k, A in enumerate(As, 1):
            _write_matrix_sdplr(f, A, k)


def read_problem_sdplr(path):
    """SDPLR-1.03 file -> (C, As, bs); sparse sections come back as symmetric CSC, low-rank ones as
    SymLowRankMatrix."""
    with open(path) as f:
        tok = f.read().split()
    p = 0

    def nxt():
        nonlocal p
        p += 1
        return tok[p - 1]

    m = int(nxt()); nblocks = int(nxt()); n = int(nxt())
    if nblocks != 1:
        raise ValueError("only single-block SDPLR problems are supported")
    bs = np.array([float(nxt()) for _ in range(m)], dtype=np.float64)
    nxt()                                             # ignored line
    mats = {}
    while p < len(tok):
        mid = int(nxt()); blk = int(nxt()); kind = nxt(); cnt = int(nxt())
        if blk != 1:
            raise ValueError("only single-block SDPLR problems are supported")
        if kind == "s":
            e = np.array([float(nxt()) for _ in range(3 * cnt)], dtype=np.float64).reshape(cnt, 3)
            i, j, v = e[:, 0].astype(np.int64) - 1, e[:, 1].astype(np.int64) - 1, e[:, 2]
            off = i != j
            mats[mid] = sps.coo_matrix((np.r_[v, v[off]], (np.r_[i, j[off]], np.r_[j, i[off]])), shape=(n, n)).tocsc()
        elif kind == "l":
            D = np.array([float(nxt()) for _ in range(cnt)], dtype=np.float64)
            B = np.array([float(nxt()) for _ in range(cnt * n)], dtype=np.float64).reshape(cnt, n).T
            mats[mid] = SymLowRankMatrix(D, B)
        else:
            raise ValueError(f"SDPLR: unknown matrix kind {kind!r}")
    return mats[0], [mats[k] for k in range(1, m + 1)], bs


def write_initial_solution(path, R, lam):
    """write_initial_solution (exps/data_utils.jl:126-152): the warm-start file of SDPLR-1.03.  R is n x r.
    (The reference omits the newline after `special lambdaupdate 0`; so does this writer.)"""
    R = np.asarray(R, np.float64)
    n, r = R.shape
    with open(path, "w") as f:
        f.write(f"dual variable {len(lam)}\n")
        for v in lam:
            f.write(jl_float(v) + "\n")
        f.write(f"primal variable 1 s {n} {r} {r}\n")
        for j in range(r):
            for i in range(n):
                f.write(jl_float(R[i, j]) + "\n")
        f.write("special majiter 0\n")
        f.write("special iter 0\n")
        f.write("special lambdaupdate 0")
        f.write("special CG 0\n")
        f.write("special curr_CG 0\n")
        f.write("special totaltime 0\n")
        f.write(f"special sigma {jl_float(1.0 / n)}\n")
        f.write("special scale 1.0\n")


def read_initial_solution(path):
    """-> (R (n x r), lambda) from an SDPLR-1.03 solution file (usable as `init_func` of sdplr)."""
    with open(path) as f:
        tok = f.read().split()
    assert tok[0] == "dual" and tok[1] == "variable"
    m = int(tok[2])
    lam = np.array(tok[3:3 + m], dtype=np.float64)
    p = 3 + m
    assert tok[p] == "primal" and tok[p + 1] == "variable"
    n, r = int(tok[p + 4]), int(tok[p + 5])
    vals = np.array(tok[p + 7:p + 7 + n * r], dtype=np.float64)
    return vals.reshape(r, n).T.copy(), lam


# ---------------------------------------------------------------------------------------------------------
# graphs
# ---------------------------------------------------------------------------------------------------------
def read_gset(path):
    """read_gset (exps/data_utils.jl:212-243): `n m` then `u v w` lines (1-based) -> symmetric CSC adjacency,
    duplicate edges summed (Julia's sparse()), self-loops removed, explicit zeros dropped."""
    with open(path) as f:
        first = f.readline().split()
        n = int(first[0])
        e = np.loadtxt(f, dtype=np.float64, ndmin=2)
    if e.size == 0:
        return sps.csc_matrix((n, n))
    u, v, w = e[:, 0].astype(np.int64) - 1, e[:, 1].astype(np.int64) - 1, e[:, 2]
    A = sps.coo_matrix((np.r_[w, w], (np.r_[u, v], np.r_[v, u])), shape=(n, n)).tocsc()
    A.setdiag(0.0)
    A.eliminate_zeros()
    A.sort_indices()
    return A


def write_gset(path, A):
    """The inverse of read_gset: one `u v w` line per upper-triangular entry."""
    A = sps.triu(sps.csc_matrix(A), k=1).tocoo()
    with open(path, "w") as f:
        f.write(f"{A.shape[0]} {A.nnz}\n")
        for u, v, w in zip(A.row, A.col, A.data):
            f.write(f"{u + 1} {v + 1} {int(w) if float(w).is_integer() else jl_float(w)}\n")


# ---------------------------------------------------------------------------------------------------------
# MATLAB v7.3 (.mat = HDF5) sparse matrices, as shipped under exps/data/*/G*.mat and read by `matread`
# (exps/data_utils.jl:7-13).  A minimal reader of the subset of HDF5 that MATLAB writes for one sparse variable:
# version-0 superblock behind the 512-byte MATLAB header, groups as link messages or symbol tables, version-1
# object headers, contiguous (or compact) dataset layouts, fixed-point / IEEE little-endian datatypes.
# ---------------------------------------------------------------------------------------------------------
class _H5:
    def __init__(self, raw):
        self.raw = raw
        sig = b"\x89HDF\r\n\x1a\n"
        self.base = raw.find(sig)
        if self.base < 0:
            raise ValueError("not an HDF5 / MATLAB v7.3 file")
        b = self.base
        if raw[b + 8] != 0:
            raise ValueError("only version-0 HDF5 superblocks (what MATLAB writes) are supported")
        self.so, self.sl = raw[b + 13], raw[b + 14]   # size of offsets / lengths
        if (self.so, self.sl) != (8, 8):
            raise ValueError("unexpected HDF5 offset/length size")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", raw, b + 16)
        base_addr, = struct.unpack_from("<Q", raw, b + 24)
        self.shift = b + base_addr if base_addr == 0 else base_addr   # file addresses are relative to the superblock
        # root group symbol table entry follows the four addresses
        ent = b + 24 + 4 * 8
        self.root_header = self._addr(struct.unpack_from("<Q", raw, ent + 8)[0])

    def _addr(self, a):
        return None if a == 0xFFFFFFFFFFFFFFFF else a + self.shift

    # -- object headers (version 1) ----------------------------------------------------------------------
    def messages(self, addr):
        raw = self.raw
        if raw[addr] != 1:
            raise ValueError("only version-1 object headers are supported")
        nmsg, = struct.unpack_from("<H", raw, addr + 2)
        hsize, = struct.unpack_from("<I", raw, addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", raw, p)
                body = p + 8
                if mtype == 0x10:                         # continuation
                    off, ln = struct.unpack_from("<QQ", raw, body)
                    blocks.append((self._addr(off), ln))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # -- groups --------------------------------------------------------------------------------------------
    def children(self, addr):
        """{name: object header address} of a group: link messages stored in the object header (what MATLAB R2006b+
        writes for small groups) or an old-style symbol table."""
        raw = self.raw
        out = {}
        for mtype, body, _ in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", raw, body)
                out.update(self._walk_btree(self._addr(btree), self._heap_data(self._addr(heap))))
            elif mtype == 0x06:                           # link message, version 1
                flags = raw[body + 1]
                p = body + 2
                ltype = 0
                if flags & 0x08:
                    ltype = raw[p]; p += 1
                if flags & 0x04:
                    p += 8                                # creation order
                if flags & 0x10:
                    p += 1                                # character set
                lsz = 1 << (flags & 3)
                ln = int.from_bytes(raw[p:p + lsz], "little"); p += lsz
                name = raw[p:p + ln].decode(); p += ln
                if ltype == 0:                            # hard link
                    out[name] = self._addr(struct.unpack_from("<Q", raw, p)[0])
        return out

    def _heap_data(self, addr):
        assert self.raw[addr:addr + 4] == b"HEAP"
        data_addr, = struct.unpack_from("<Q", self.raw, addr + 8 + 16)
        return self._addr(data_addr)

    def _walk_btree(self, addr, heap):
        raw = self.raw
        out = {}
        assert raw[addr:addr + 4] == b"TREE"
        level = raw[addr + 5]
        used, = struct.unpack_from("<H", raw, addr + 6)
        p = addr + 8 + 16          # skip sibling addresses
        p += 8                     # key 0
        for _ in range(used):
            child, = struct.unpack_from("<Q", raw, p)
            p += 8 + 8             # child + next key
            caddr = self._addr(child)
            if level > 0:
                out.update(self._walk_btree(caddr, heap))
            else:
                assert raw[caddr:caddr + 4] == b"SNOD"
                nsym, = struct.unpack_from("<H", raw, caddr + 6)
                e = caddr + 8
                for _ in range(nsym):
                    name_off, hdr = struct.unpack_from("<QQ", raw, e)
                    nm_end = raw.index(b"\x00", heap + name_off)
                    out[raw[heap + name_off:nm_end].decode()] = self._addr(hdr)
                    e += 40
        return out

    # -- datasets ------------------------------------------------------------------------------------------
    def dataset(self, addr):
        raw = self.raw
        shape = dtype = data = None
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x01:                             # dataspace
                ver, rank, flags = raw[body], raw[body + 1], raw[body + 2]
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<" + "Q" * rank, raw, off)
            elif mtype == 0x03:                           # datatype
                cls = raw[body] & 0x0F
                bits0 = raw[body + 1]
                size, = struct.unpack_from("<I", raw, body + 4)
                if bits0 & 1:
                    raise ValueError("big-endian HDF5 data is not supported")
                if cls == 0:
                    dtype = np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
                elif cls == 1:
                    dtype = np.dtype("<f" + str(size))
                else:
                    raise ValueError(f"unsupported HDF5 datatype class {cls}")
            elif mtype == 0x08:                           # data layout
                ver = raw[body]
                if ver != 3:
                    raise ValueError("only version-3 data layout messages are supported")
                lclass = raw[body + 1]
                if lclass == 1:                           # contiguous
                    a, ln = struct.unpack_from("<QQ", raw, body + 2)
                    data = (self._addr(a), ln)
                elif lclass == 0:                         # compact
                    ln, = struct.unpack_from("<H", raw, body + 2)
                    data = (body + 4, ln)
                else:
                    raise ValueError("chunked / compressed HDF5 datasets are not supported (MATLAB writes small sparse "
                                     "matrices contiguously)")
        if shape is None or dtype is None or data is None:
            raise ValueError("incomplete HDF5 dataset header")
        count = int(np.prod(shape)) if len(shape) else 1
        if data[0] is None or count == 0:
            return np.zeros(shape, dtype)
        return np.frombuffer(raw, dtype=dtype, count=count, offset=data[0]).reshape(shape)


def read_mat_sparse(path, name=None):
    """`matread(path)[name]` for a MATLAB v7.3 file holding a sparse matrix (group with datasets `data`, `ir`, `jc`):
    -> scipy CSC.  name=None takes the first sparse variable (the reference's files hold one, `A`)."""
    with open(path, "rb") as f:
        raw = f.read()
    if not raw.startswith(b"MATLAB 7.3 MAT-file"):
        raise ValueError("not a MATLAB v7.3 MAT-file (older formats: scipy.io.loadmat)")
    h5 = _H5(raw)
    top = h5.children(h5.root_header)
    names = [name] if name is not None else sorted(k for k in top if not k.startswith("#"))
    for nm in names:
        if nm not in top:
            raise KeyError(nm)
        kids = h5.children(top[nm])
        if {"ir", "jc"} <= set(kids):
            jc = h5.dataset(kids["jc"]).astype(np.int64).reshape(-1)
            ir = h5.dataset(kids["ir"]).astype(np.int64).reshape(-1)
            data = h5.dataset(kids["data"]).astype(np.float64).reshape(-1) if "data" in kids else np.ones(ir.size)
            ncols = jc.size - 1
            nrows = int(ir.max()) + 1 if ir.size else ncols
            nrows = max(nrows, ncols)      # the shipped graphs are square; MATLAB_sparse attribute is not parsed
            return sps.csc_matrix((data, ir, jc), shape=(nrows, ncols))
    raise ValueError("no sparse variable found in " + os.path.basename(path))


def read_graph(path, name=None):
    """read_graph (exps/data_utils.jl:1-13): adjacency matrix from a `.mat` (v7.3) or Gset text file."""
    with open(path, "rb") as f:
        head = f.read(19)
    if head == b"MATLAB 7.3 MAT-file":
        return read_mat_sparse(path, name)
    return read_gset(path)
