"""ctypes wrapper of oracle/liboracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  Never by the product package.

`OracleEngine` exposes the same engine interface as sdplrplus.jl_b200's
B200Engine so the same `_sdplr` driver and the same tests can run against
either implementation.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_p_i64 = C.POINTER(C.c_int64)
_p_f64 = C.POINTER(C.c_double)
_p_u8 = C.POINTER(C.c_uint8)
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "sdplr_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    V = C.c_void_p
    sig = {
        "orc_create": ([C.c_int64, C.c_int64], V),
        "orc_destroy": ([V], None),
        "orc_set_threads": ([C.c_int], None),
        "orc_max_threads": ([], C.c_int),
        "orc_preprocess": ([V, C.c_int64, _p_i64, _p_i64, _p_i64, _p_f64, _p_i64], C.c_int),
        "orc_set_maps": ([V, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _p_i64, _p_i64, _p_i64, _p_i64, _p_f64, _p_f64,
                          _p_i64, _p_i64, _p_i64, _p_i64], C.c_int),
        "orc_pattern_sizes": ([V, _p_i64, _p_i64, _p_i64], None),
        "orc_pattern_export": ([V, _p_i64, _p_i64, _p_i64, _p_i64, _p_f64, _p_f64, _p_i64, _p_i64, _p_i64], None),
        "orc_add_symlowrank": ([V, C.c_int64, C.c_int64, _p_f64, _p_f64], C.c_int),
        "orc_set_problem": ([V, _p_f64, _p_u8], None),
        "orc_init_vars": ([V, C.c_int64, _p_f64, _p_f64, C.c_double], None),
        "orc_ptr": ([V, C.c_int], _p_f64),
        "orc_get_sigma": ([V], C.c_double), "orc_set_sigma": ([V, C.c_double], None),
        "orc_get_obj": ([V], C.c_double), "orc_set_obj": ([V, C.c_double], None),
        "orc_A_uu": ([V, _p_f64, C.c_int64, _p_f64], None),
        "orc_A_uv": ([V, _p_f64, C.c_int64, _p_f64, _p_f64], None),
        "orc_f": ([V], C.c_double),
        "orc_copy2y": ([V], None),
        "orc_At_preprocess": ([V], None),
        "orc_At_left": ([V, _p_f64, C.c_int64, _p_f64], None),
        "orc_At_right": ([V, _p_f64, C.c_int64, _p_f64], None),
        "orc_g": ([V], None),
        "orc_fg": ([V, C.c_double, C.c_double, _p_f64], None),
        "orc_dot": ([_p_f64, _p_f64, C.c_int64], C.c_double),
        "orc_nrm2": ([_p_f64, C.c_int64], C.c_double),
        "orc_axpy": ([C.c_double, _p_f64, _p_f64, C.c_int64], None),
        "orc_scal": ([C.c_double, _p_f64, C.c_int64], None),
        "orc_biquadratic": ([V, _p_f64], None),
        "orc_pick_alpha": ([_p_f64, C.c_double, _p_f64, _p_f64], C.c_int),
        "orc_commit_step": ([V, C.c_double], None),
        "orc_linesearch_passes": ([V, _p_f64], None),
        "orc_linesearch": ([V, _p_f64, C.c_double, _p_f64, _p_f64, _p_f64], C.c_int),
        "orc_eval_AL": ([V, C.c_double], C.c_double),
        "orc_linesearch_armijo": ([V, _p_f64, C.c_double, _p_f64, _p_f64], C.c_int),
        "orc_lbfgs_init": ([V, C.c_int64], None),
        "orc_lbfgs_clear": ([V], None),
        "orc_lbfgs_dir": ([V, _p_f64, _p_f64, C.c_int], None),
        "orc_lbfgs_update": ([V, _p_f64, _p_f64, C.c_double], None),
        "orc_lbfgs_ptr": ([V, C.c_int, C.c_int64], _p_f64),
        "orc_lbfgs_rho": ([V, C.c_int64], C.c_double),
        "orc_lbfgs_latest": ([V], C.c_int64),
        "orc_tridiag_mineig": ([_p_f64, _p_f64, C.c_int64], C.c_double),
        "orc_lanczos": ([V, C.c_int64, _p_f64, _p_i64, _p_f64, _p_f64], C.c_double),
        "orc_dual_obj": ([V, C.c_double, C.c_int64, _p_f64, _p_f64, _p_i64], C.c_double),
        "orc_dual_update": ([V], None),
        "orc_copy2y_lambda": ([V], None),
        "orc_dual_value": ([V, C.c_double, C.c_double], C.c_double),
        "orc_dimacs_errors": ([V, C.c_double, C.c_double, C.c_double, _p_f64], None),
        "orc_symlowrank_norm": ([C.c_int64, C.c_int64, _p_f64, _p_f64, C.c_int], C.c_double),
    }
    for name, (args, res) in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    _lib = lib
    return lib


def _f(a):
    return a.ctypes.data_as(_p_f64)


def _i(a):
    return a.ctypes.data_as(_p_i64)


PTR = dict(Rt=0, Gt=1, lam=2, lam_ub=3, y=4, pvio_raw=5, pvio_lb=6, pvio=7, A_RD=8, A_DD=9, b=10, triuS=11, S=12)


class Oracle:
    """Thin object wrapper of one orc_ctx."""

    def __init__(self, asm, b=None, is_ineq=None, maps=None):
        """asm: sdplrplus.jl_b200.types.AssembledSparse (1-based triplets)."""
        self.lib = load()
        self.n, self.m = int(asm.n), int(asm.m)
        self.ctx = self.lib.orc_create(self.n, self.m)
        self.nA = len(asm.gids)
        self.r = 0
        if maps is None:
            I = np.ascontiguousarray(asm.I, np.int64); J = np.ascontiguousarray(asm.J, np.int64)
            V = np.ascontiguousarray(asm.V, np.float64); off = np.ascontiguousarray(asm.mat_off, np.int64)
            g = np.ascontiguousarray(asm.gids, np.int64)
            self.preprocess_rc = self.lib.orc_preprocess(self.ctx, self.nA, _i(off), _i(I), _i(J), _f(V), _i(g))
            if self.preprocess_rc < 0:
                raise ValueError("oracle: coordinate out of range")
        else:
            g = np.ascontiguousarray(asm.gids, np.int64)
            ks = ["triu_colptr", "triu_rowval", "matptr", "nzind", "nzval_one", "nzval_two", "full_colptr", "full_rowval", "mapped"]
            a = {k: np.ascontiguousarray(maps[k]) for k in ks}
            self.lib.orc_set_maps(self.ctx, self.nA, a["triu_rowval"].size, a["full_rowval"].size, a["nzind"].size,
                                  _i(a["triu_colptr"]), _i(a["triu_rowval"]), _i(a["matptr"]), _i(a["nzind"]),
                                  _f(a["nzval_one"]), _f(a["nzval_two"]), _i(a["full_colptr"]), _i(a["full_rowval"]),
                                  _i(a["mapped"]), _i(g))
            self.preprocess_rc = 0
        for gid1, A in asm.lowrank:
            B = np.asfortranarray(A.B, dtype=np.float64); D = np.ascontiguousarray(A.D, np.float64)
            self.lib.orc_add_symlowrank(self.ctx, gid1, B.shape[1], _f(B), _f(D))
        if b is not None:
            self.set_problem(b, is_ineq)

    def __del__(self):
        try:
            if self.ctx:
                self.lib.orc_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def set_problem(self, b, is_ineq=None):
        b = np.ascontiguousarray(b, np.float64)
        q = None if is_ineq is None else np.ascontiguousarray(is_ineq, np.uint8)
        self.lib.orc_set_problem(self.ctx, _f(b), None if q is None else q.ctypes.data_as(_p_u8))

    def pattern_sizes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.orc_pattern_sizes(self.ctx, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def pattern_export(self):
        nnzT, nnzF, Ec = self.pattern_sizes()
        out = {"triu_colptr": np.zeros(self.n + 1, np.int64), "triu_rowval": np.zeros(nnzT, np.int64),
               "matptr": np.zeros(self.nA + 1, np.int64), "nzind": np.zeros(Ec, np.int64),
               "nzval_one": np.zeros(Ec), "nzval_two": np.zeros(Ec), "full_colptr": np.zeros(self.n + 1, np.int64),
               "full_rowval": np.zeros(nnzF, np.int64), "mapped": np.zeros(nnzF, np.int64)}
        self.lib.orc_pattern_export(self.ctx, _i(out["triu_colptr"]), _i(out["triu_rowval"]), _i(out["matptr"]),
                                    _i(out["nzind"]), _f(out["nzval_one"]), _f(out["nzval_two"]), _i(out["full_colptr"]),
                                    _i(out["full_rowval"]), _i(out["mapped"]))
        return out

    def view(self, name, length=None):
        """numpy view of an oracle-owned array (no copy)."""
        which = PTR[name]
        if length is None:
            length = {0: self.n * self.r, 1: self.n * self.r, 2: self.m, 3: self.m, 4: self.m + 1, 5: self.m + 1,
                      6: self.m, 7: self.m, 8: self.m + 1, 9: self.m + 1, 10: self.m}[which]
        p = self.lib.orc_ptr(self.ctx, which)
        return np.ctypeslib.as_array(p, shape=(length,))

    def init_vars(self, r, Rt0, lam0, sigma0):
        Rt0 = np.ascontiguousarray(Rt0, np.float64)
        assert Rt0.size == self.n * r
        lam0 = np.ascontiguousarray(lam0, np.float64)
        self.lib.orc_init_vars(self.ctx, r, _f(Rt0), _f(lam0), float(sigma0))
        self.r = int(r)

    sigma = property(lambda s: s.lib.orc_get_sigma(s.ctx), lambda s, v: s.lib.orc_set_sigma(s.ctx, float(v)))
    obj = property(lambda s: s.lib.orc_get_obj(s.ctx))

    def A_uu(self, Ut):
        Ut = np.ascontiguousarray(Ut, np.float64); out = np.zeros(self.m + 1)
        self.lib.orc_A_uu(self.ctx, _f(out), Ut.size // self.n, _f(Ut))
        return out

    def A_uv(self, Ut, Vt):
        Ut = np.ascontiguousarray(Ut, np.float64); Vt = np.ascontiguousarray(Vt, np.float64); out = np.zeros(self.m + 1)
        self.lib.orc_A_uv(self.ctx, _f(out), Ut.size // self.n, _f(Ut), _f(Vt))
        return out

    def At_preprocess(self, y=None):
        if y is not None:
            self.view("y")[:] = y
        self.lib.orc_At_preprocess(self.ctx)

    def At_left(self, X):
        X = np.ascontiguousarray(X, np.float64); Y = np.zeros_like(X)
        self.lib.orc_At_left(self.ctx, _f(Y), X.size // self.n, _f(X))
        return Y

    def At_right(self, x):
        x = np.asarray(x, np.float64); one = x.ndim == 1
        x2 = np.asfortranarray(x.reshape(-1, 1) if one else x)
        y = np.zeros(x2.shape, order="F")
        self.lib.orc_At_right(self.ctx, _f(y), x2.shape[1], _f(x2))
        return y[:, 0] if one else y

    def f(self):
        return self.lib.orc_f(self.ctx)

    def g(self):
        self.lib.orc_g(self.ctx)

    def lanczos(self, q, v0):
        q = int(max(1, min(q, self.n - 1)))
        v0 = np.ascontiguousarray(v0, np.float64); a = np.zeros(q); b = np.zeros(q); it = C.c_int64()
        lam = self.lib.orc_lanczos(self.ctx, q, _f(v0), C.byref(it), _f(a), _f(b))
        return lam, a, b, it.value


def S_eigval(o, nevs=1, ncv=None, tol=0.0, maxiter=10 ** 6, v0=None):
    """SDP_S_eigval (src/coreop.jl:351-374) on the oracle's current S.

    The arithmetic lives in a third-party dependency that is not under /root/reference: GenericArpack.jl
    (Project.toml compat "0.2"), a Julia port of ARPACK's dsaupd/dseupd, i.e. the implicitly restarted Lanczos method
    (Lehoucq-Sorensen-Yang) with exact shifts, Ritz estimate rule |beta_m y_m,i| <= tol*max(eps^(2/3), |theta_i|), and
    tol = 0 meaning machine precision.  scipy.sparse.linalg.eigsh binds ARPACK itself, so the same published algorithm
    is applied to the same operator x -> S*x + x with the same (which, ncv, tol); small problems (n <= 3, where ARPACK
    needs nev < ncv <= n) fall back to the dense eigenvalues.  Parity is to solver tolerance (the start vector of
    GenericArpack is random and not reproducible)."""
    import scipy.sparse.linalg as sla
    n = o.n
    ncv = min(100, n) if ncv is None else min(int(ncv), n)
    if n <= 3 or nevs >= n - 1:
        S = np.column_stack([o.At_right(e) for e in np.eye(n)])
        return np.sort(np.linalg.eigvalsh(0.5 * (S + S.T)))[:nevs]
    op = sla.LinearOperator((n, n), matvec=lambda x: o.At_right(np.asarray(x, np.float64).reshape(-1)) + np.asarray(x).reshape(-1),
                            dtype=np.float64)
    ev = sla.eigsh(op, k=int(nevs), which="SA", ncv=max(ncv, int(nevs) + 1), tol=float(tol), maxiter=int(maxiter), v0=v0,
                   return_eigenvectors=False)
    return np.sort(np.real(ev)) - 1.0


class OracleEngine:
    """Engine interface of sdplrplus.jl_b200.solver (see B200Engine) on the CPU oracle."""

    def __init__(self, data, asm=None, maps=None):
        if asm is None:
            from sdplrplus.jl_b200.types import assemble_sparse
            asm = assemble_sparse(data)
        self.data = data
        self.o = Oracle(asm, data.b, data.constraint_types.astype(np.uint8) if data.has_inequalities else None, maps=maps)
        self.lib = self.o.lib
        self.n, self.m = data.n, data.m
        self.Dt = None

    def init_vars(self, r, Rt0, lambda0, sigma0, numlbfgsvecs):
        self.o.init_vars(r, Rt0, lambda0, sigma0)
        self.lib.orc_lbfgs_init(self.o.ctx, int(numlbfgsvecs))
        self.r = r
        self.Dt = np.zeros(self.n * r)

    sigma = property(lambda s: s.o.sigma, lambda s, v: setattr(s.o, "sigma", v))

    def get_R(self):
        return self.o.view("Rt").reshape(self.n, self.r).copy()

    def get_G(self):
        return self.o.view("Gt").reshape(self.n, self.r).copy()

    def get_D(self):
        return self.Dt.reshape(self.n, self.r).copy()

    def get_lambda(self):
        return self.o.view("lam").copy()

    def get_y(self):
        return self.o.view("y").copy()

    def get_pvio_raw(self):
        return self.o.view("pvio_raw").copy()

    def set_R(self, Rt):
        self.o.view("Rt")[:] = np.ascontiguousarray(Rt, np.float64).reshape(-1)

    def set_D(self, Dt):
        self.Dt[:] = np.ascontiguousarray(Dt, np.float64).reshape(-1)

    def set_lambda(self, lam):
        self.o.view("lam")[:] = lam

    def fg(self):
        out = np.zeros(3)
        self.lib.orc_fg(self.o.ctx, 1.0, 1.0, _f(out))
        return float(out[0]), self.o.obj, float(out[1]) ** 2, float(out[2]) ** 2

    def f(self):
        L = self.lib.orc_f(self.o.ctx)
        return L, self.o.obj

    def g(self):
        self.lib.orc_g(self.o.ctx)
        N = self.n * self.r
        gn = self.lib.orc_nrm2(self.lib.orc_ptr(self.o.ctx, 1), N)
        pv = np.maximum(self.o.view("pvio_raw")[: self.m], self.o.view("pvio_lb"))
        return gn * gn, float(np.dot(pv, pv))

    def lbfgs_dir(self):
        self.lib.orc_lbfgs_dir(self.o.ctx, _f(self.Dt), self.lib.orc_ptr(self.o.ctx, 1), 1)
        return self.lib.orc_dot(_f(self.Dt), self.lib.orc_ptr(self.o.ctx, 1), self.n * self.r)

    def use_gradient_direction(self):
        G = self.o.view("Gt")
        G *= -1.0
        self.Dt[:] = G

    def linesearch_coeffs(self):
        self.lib.orc_linesearch_passes(self.o.ctx, _f(self.Dt))
        bq = np.zeros(5)
        self.lib.orc_biquadratic(self.o.ctx, _f(bq))
        return bq

    def armijo_eval(self, alphas):
        L = np.array([self.lib.orc_eval_AL(self.o.ctx, float(a)) for a in np.atleast_1d(alphas)])
        slope = self.o.view("A_RD")[self.m] + float(np.dot(self.o.view("y")[: self.m], self.o.view("A_RD")[: self.m]))
        return L, slope

    def step(self, alpha):
        self.lib.orc_commit_step(self.o.ctx, float(alpha))
        self.lib.orc_axpy(float(alpha), _f(self.Dt), self.lib.orc_ptr(self.o.ctx, 0), self.n * self.r)
        return self.o.obj

    def lbfgs_update(self, alpha):
        self.lib.orc_lbfgs_update(self.o.ctx, _f(self.Dt), self.lib.orc_ptr(self.o.ctx, 1), float(alpha))

    def lbfgs_clear(self):
        self.lib.orc_lbfgs_clear(self.o.ctx)

    def dual_obj(self, trace_bound, it, v0=None, seed=0):
        if v0 is None:
            v0 = np.random.default_rng(seed).standard_normal(self.n)
        v0 = np.ascontiguousarray(v0, np.float64)
        lam = C.c_double(); q = C.c_int64()
        d = self.lib.orc_dual_obj(self.o.ctx, float(trace_bound), int(it), _f(v0), C.byref(lam), C.byref(q))
        return d, lam.value, q.value

    def dual_obj_highprecision(self, trace_bound, v0=None, seed=0):
        """dual_obj(...; highprecision=true) (src/coreop.jl:376-415)"""
        self.lib.orc_copy2y(self.o.ctx)
        self.lib.orc_At_preprocess(self.o.ctx)
        lam = float(S_eigval(self.o, 1, min(100, self.n), 1e-6, 10 ** 6, v0)[0])
        return self.lib.orc_dual_value(self.o.ctx, float(trace_bound), lam), lam, 0

    def dimacs_errors(self, normb, normC, v0=None, seed=0):
        """DIMACS_errors (src/coreop.jl:426-453)"""
        self.lib.orc_copy2y_lambda(self.o.ctx)
        self.lib.orc_At_preprocess(self.o.ctx)
        lam = float(S_eigval(self.o, 1, min(100, self.n), 0.0, 10 ** 6, v0)[0])
        errs = np.zeros(6)
        self.lib.orc_dimacs_errors(self.o.ctx, float(normb), float(normC), lam, _f(errs))
        return errs

    def dual_update(self):
        self.lib.orc_dual_update(self.o.ctx)

    def close(self):
        pass
