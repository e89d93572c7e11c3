"""ctypes front-end of the C oracle (``oracle/fos_oracle.c``).

TEST INFRASTRUCTURE ONLY -- see the header of ``fos_oracle.c``.  Nothing under
``firstordersolvers.jl_b200/`` may import this module; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

The class mirrors the reference objects it restates:

* ``OracleConic``       -- ``FOSMathProgModel`` + ``HSDE`` (src/problemforms/HSDE/HSDE.jl:7-29)
* ``OracleFeasibility`` -- ``Feasibility(AffinePlusLinear, ConeProduct, n)``
  (src/problemforms/Feasibility/Feasibility.jl:2-6)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import scipy.sparse as sp

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libfos_oracle.so"
# build variants of the one C file (see its header): "" the restatement proper, "hp" long-double reductions
# (the "exact" yardstick of the parity tests), "mt" OpenMP over the sparse products (timing baseline only)
VARIANTS = {"": "libfos_oracle.so", "hp": "libfos_oracle_hp.so", "mt": "libfos_oracle_mt.so"}

CONE_CODES = {"Free": 0, "Zero": 1, "NonNeg": 2, "NonPos": 3, "SOC": 4, "SOCRotated": 5, "SDP": 6,
              "ExpPrimal": 7, "ExpDual": 8}
ALG_CODES = {"GAP": 0, "GAPA": 1, "FISTA": 2, "Dykstra": 3, "GAPP": 4}
STATUS_NAMES = {0: "Continue", 1: "Optimal", 2: "Unbounded", 3: "Infeasible", 4: "Indeterminate"}
REC_LEN = 10
REC_FIELDS = ("i", "p", "d", "g", "ctx", "bty", "kappa", "tau", "cgiter", "status")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False, variant: str = "") -> Path:
    """Compile ``libfos_oracle[_hp|_mt].so`` with the committed Makefile (gcc only)."""
    src = HERE / "fos_oracle.c"
    path = HERE / VARIANTS[variant]
    if force or not path.exists() or path.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-s", "-C", str(HERE), VARIANTS[variant]], check=True)
    return path


def build_all(force: bool = False):
    """All three variants; the threaded one is optional (needs libgomp) and its absence is reported, not fatal."""
    out = {}
    for v in VARIANTS:
        try:
            out[v] = build(force, v)
        except subprocess.CalledProcessError:
            if v != "mt":
                raise
            out[v] = None
    return out


_libs = {}


def lib(variant: str = ""):
    if variant in _libs:
        return _libs[variant]
    path = build(variant=variant)
    L = C.CDLL(str(path))
    L.fosor_create_conic.restype = C.c_void_p
    L.fosor_create_conic.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int64, _dp, _dp,
                                     C.c_int64, _i32p, _ip, C.c_int64, _i32p, _ip]
    L.fosor_create_feasibility.restype = C.c_void_p
    L.fosor_create_feasibility.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int64, _dp, _dp,
                                           C.c_int64, C.c_int32, C.c_int64, _i32p, _ip]
    L.fosor_destroy.argtypes = [C.c_void_p]
    L.fosor_set_box.restype = C.c_int32
    L.fosor_set_box.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double]
    L.fosor_set_linesearch.restype = C.c_int32
    L.fosor_set_linesearch.argtypes = [C.c_void_p, C.c_int64]
    L.fosor_set_direct.restype = C.c_int32
    L.fosor_set_direct.argtypes = [C.c_void_p, C.c_int32]
    L.fosor_iterate_length.restype = C.c_int64
    L.fosor_iterate_length.argtypes = [C.c_void_p]
    L.fosor_set_algorithm.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double,
                                      C.c_int64]
    L.fosor_set_iterate.argtypes = [C.c_void_p, _dp]
    L.fosor_get_iterate.argtypes = [C.c_void_p, _dp]
    L.fosor_get_state.argtypes = [C.c_void_p, C.c_int32, _dp]
    L.fosor_get_s1_calls.restype = C.c_int64
    L.fosor_get_s1_calls.argtypes = [C.c_void_p]
    L.fosor_get_cgiter.restype = C.c_int64
    L.fosor_get_cgiter.argtypes = [C.c_void_p]
    L.fosor_get_alpha12.restype = C.c_double
    L.fosor_get_alpha12.argtypes = [C.c_void_p]
    L.fosor_get_fista_t.restype = C.c_double
    L.fosor_get_fista_t.argtypes = [C.c_void_p]
    L.fosor_get_cg_warned.restype = C.c_int32
    L.fosor_get_cg_warned.argtypes = [C.c_void_p]
    L.fosor_run.restype = C.c_int64
    L.fosor_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, _dp, C.c_int64, _ip, _dp,
                            _i32p]
    L.fosor_finish.argtypes = [C.c_void_p, _dp, _dp, _ip, _i32p]
    L.fosor_solve.restype = C.c_int64
    L.fosor_solve.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_double, _dp, _dp, C.c_int64, _ip, _i32p]
    L.fosor_populate_solution.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    L.fosor_q_mul.argtypes = [C.c_void_p, _dp, _dp, C.c_int32]
    L.fosor_kkt_mul.argtypes = [C.c_void_p, _dp, _dp]
    L.fosor_affine_prox.argtypes = [C.c_void_p, _dp, _dp]
    L.fosor_cone_prox.argtypes = [C.c_void_p, _dp, _dp]
    L.fosor_a_mul.argtypes = [C.c_void_p, _dp, _dp, C.c_int32]
    L.fosor_hsdematrix_prox.argtypes = [C.c_void_p, _dp, _dp]
    L.fosor_cg_csc.restype = C.c_int64
    L.fosor_cg_csc.argtypes = [C.c_int64, _ip, _ip, _dp, C.c_int64, _dp, _dp, C.c_double, C.c_int64]
    L.fosor_prox_cone.restype = C.c_int32
    L.fosor_prox_cone.argtypes = [C.c_int32, C.c_int32, _dp, _dp, C.c_int64]
    L.fosor_set_state.argtypes = [C.c_void_p, C.c_int32, _dp]
    L.fosor_set_scalar.argtypes = [C.c_void_p, C.c_int32, C.c_double]
    L.fosor_variant.restype = C.c_int32
    L.fosor_threads.restype = C.c_int32
    L.fosor_create_conic_borrowed.restype = C.c_void_p
    L.fosor_create_conic_borrowed.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, _dp, C.c_int64, _i32p, _ip,
                                              C.c_int64, _i32p, _ip]
    L.fosor_gen_dense_csc.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_double, _ip, _ip, _dp]
    L.fosor_csc_mul.argtypes = [C.c_int64, C.c_int64, _ip, _ip, _dp, _dp, _dp, C.c_int32]
    _libs[variant] = L
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _i32(a):
    return a.ctypes.data_as(_i32p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _cones(cones):
    """[(name, length), ...] -> (types int32[], lens int64[])"""
    t = np.array([CONE_CODES[c[0]] for c in cones], dtype=np.int32)
    ln = np.array([int(c[1]) for c in cones], dtype=np.int64)
    return t, ln


def _csc(A):
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sort_indices()
    return (A, np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int64),
            np.ascontiguousarray(A.data, dtype=np.float64))


def records_to_history(rec: np.ndarray) -> dict:
    """(k, REC_LEN) record array -> dict of arrays keyed like the reference's MVHistory."""
    rec = np.asarray(rec).reshape(-1, REC_LEN)
    return {name: rec[:, j].copy() for j, name in enumerate(REC_FIELDS)}


class _OracleBase:
    def __init__(self, variant=""):
        self._h = None
        self.N = 0
        self.variant = variant
        self._L = lib(variant)

    def __del__(self):
        try:
            if self._h:
                self._L.fosor_destroy(self._h)
                self._h = None
        except Exception:
            pass

    _STATE = {"x": 0, "tmp1": 1, "tmp2": 2, "xinit": 3, "rhs": 4, "fista_y": 6, "dykstra_p": 7, "dykstra_q": 8}

    def set_state(self, which, z):
        """Restores a persistent vector ("xinit" also clears S1's first-run flag)."""
        z = _f64(z)
        assert z.shape == (self.N,)
        self._L.fosor_set_state(self._h, self._STATE[which], _d(z))

    def set_scalar(self, which, value):
        """which: "s1_calls", "alpha12" or "fista_t"."""
        self._L.fosor_set_scalar(self._h, {"s1_calls": 0, "alpha12": 2, "fista_t": 3}[which], float(value))

    # -- algorithm (a20) -------------------------------------------------------------
    def set_algorithm(self, name, alpha=0.8, alpha1=1.8, alpha2=1.8, beta=0.0, iproj=100):
        self._L.fosor_set_algorithm(self._h, ALG_CODES[name], alpha, alpha1, alpha2, beta, iproj)

    def set_box(self, start, length, lo, hi):
        """IndBox(lo, hi) on entries [start, start+length) of the Feasibility iterate (0-based)."""
        if self._L.fosor_set_box(self._h, int(start), int(length), float(lo), float(hi)) != 0:
            raise ValueError("bad box")

    def set_linesearch(self, lsinterval):
        """LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl:19-24); 0 removes the wrapper."""
        if self._L.fosor_set_linesearch(self._h, int(lsinterval)) != 0:
            raise ValueError("bad lsinterval")

    def set_iterate(self, z):
        z = _f64(z)
        assert z.shape == (self.N,)
        self._L.fosor_set_iterate(self._h, _d(z))

    def get_iterate(self):
        z = np.empty(self.N)
        self._L.fosor_get_iterate(self._h, _d(z))
        return z

    def get_state(self, which):
        idx = {"x": 0, "tmp1": 1, "tmp2": 2, "xinit": 3, "rhs": 4, "fista_y": 6, "dykstra_p": 7, "dykstra_q": 8}[which]
        z = np.empty(self.N)
        self._L.fosor_get_state(self._h, idx, _d(z))
        return z

    @property
    def s1_calls(self):
        return self._L.fosor_get_s1_calls(self._h)

    @property
    def cgiter(self):
        return self._L.fosor_get_cgiter(self._h)

    @property
    def alpha12(self):
        return self._L.fosor_get_alpha12(self._h)

    @property
    def fista_t(self):
        return self._L.fosor_get_fista_t(self._h)

    def run(self, i_start, n_iters, checki=100, eps=1e-5, trace=False):
        """Iterations i_start..i_start+n_iters-1 of solverwrapper.jl:23-29."""
        cap = n_iters // max(checki, 1) + 2
        hist = np.zeros((cap, REC_LEN))
        hl = C.c_int64(0)
        st = C.c_int32(0)
        tr = np.zeros((n_iters, self.N)) if trace else None
        done = self._L.fosor_run(self._h, i_start, n_iters, checki, eps, _d(hist), cap, C.byref(hl),
                               _d(tr) if trace else None, C.byref(st))
        out = {"done": int(done), "status": STATUS_NAMES[st.value], "history": records_to_history(hist[:hl.value])}
        if trace:
            out["trace"] = tr[:done]
        return out

    def finish(self):
        """getsol + forced final check (solverwrapper.jl:31-34)."""
        guess = np.empty(self.N)
        hist = np.zeros((1, REC_LEN))
        hl = C.c_int64(0)
        st = C.c_int32(0)
        self._L.fosor_finish(self._h, _d(guess), _d(hist), C.byref(hl), C.byref(st))
        return guess, records_to_history(hist[:hl.value]), STATUS_NAMES[st.value]

    def solve(self, max_iters=10000, checki=100, eps=1e-5):
        """solve!(model) (solverwrapper.jl:2-17) on the current iterate."""
        cap = max_iters // max(checki, 1) + 2
        hist = np.zeros((cap, REC_LEN))
        hl = C.c_int64(0)
        st = C.c_int32(0)
        guess = np.empty(self.N)
        done = self._L.fosor_solve(self._h, max_iters, checki, eps, _d(guess), _d(hist), cap, C.byref(hl),
                                 C.byref(st))
        status = STATUS_NAMES[st.value]
        if status == "Continue":
            status = "Indeterminate"  # HSDE.jl:57-59 / Feasibility.jl:61-64
        return {"iterations": int(done), "status": status, "guess": guess,
                "history": records_to_history(hist[:hl.value])}

    # -- unit-level ------------------------------------------------------------------
    def kkt_mul(self, x):
        x = _f64(x)
        y = np.empty(self.N)
        self._L.fosor_kkt_mul(self._h, _d(x), _d(y))
        return y

    def affine_prox(self, x):
        x = _f64(x)
        y = np.empty(self.N)
        self._L.fosor_affine_prox(self._h, _d(x), _d(y))
        return y

    def cone_prox(self, x):
        x = _f64(x)
        y = np.empty(self.N)
        self._L.fosor_cone_prox(self._h, _d(x), _d(y))
        return y


class OracleConic(_OracleBase):
    """HSDE conic model: minimise c'x s.t. b - A x in K1, x in K2 (MathProgBase convention)."""

    def __init__(self, c, A, b, constr_cones, var_cones, direct=False, variant=""):
        super().__init__(variant)
        A, colptr, rowval, nzval = _csc(A)
        self.m, self.n = A.shape
        self.c = _f64(c)
        self.b = _f64(b)
        t1, l1 = _cones(constr_cones)
        t2, l2 = _cones(var_cones)
        self._h = self._L.fosor_create_conic(self.m, self.n, _i(colptr), _i(rowval), _d(nzval), 0, _d(self.b),
                                           _d(self.c), len(t1), _i32(t1), _i(l1), len(t2), _i32(t2), _i(l2))
        if not self._h:
            raise ValueError("cones do not cover 1:m / 1:n (cones.jl:66-72)")
        self.N = self._L.fosor_iterate_length(self._h)
        self.l = self.m + self.n + 1
        if direct:  # HSDE(model, direct=true): S1 = IndAffine([Q -I], 0)  (HSDE.jl:10-15)
            self.set_direct(True)

    def set_direct(self, on=True):
        if self._L.fosor_set_direct(self._h, 1 if on else 0) != 0:
            raise RuntimeError("factorisation of I + Q Q' failed")

    def initial_value(self):
        """HSDE_getinitialvalue (HSDE.jl:40-47)."""
        z = np.zeros(self.N)
        z[self.l - 1] = 1.0
        z[2 * self.l - 1] = 1.0
        return z

    def q_mul(self, B, transpose=False):
        B = _f64(B)
        Y = np.empty(self.l)
        self._L.fosor_q_mul(self._h, _d(B), _d(Y), 1 if transpose else 0)
        return Y

    def a_mul(self, x, transpose=False):
        x = _f64(x)
        y = np.empty(self.n if transpose else self.m)
        self._L.fosor_a_mul(self._h, _d(x), _d(y), 1 if transpose else 0)
        return y

    def hsdematrix_prox(self, x):
        x = _f64(x)
        y = np.empty(self.N)
        self._L.fosor_hsdematrix_prox(self._h, _d(x), _d(y))
        return y

    def populate_solution(self, guess):
        guess = _f64(guess)
        x = np.empty(self.n)
        y = np.empty(self.m)
        s = np.empty(self.m)
        self._L.fosor_populate_solution(self._h, _d(guess), _d(x), _d(y), _d(s))
        return x, y, s


class OracleFeasibility(_OracleBase):
    """Feasibility(S1=AffinePlusLinear(A,b,q,beta), S2=ConeProduct(cones), n=an+am)."""

    def __init__(self, A, b, q, beta, cones, decreasing_accuracy=False, variant=""):
        super().__init__(variant)
        A, colptr, rowval, nzval = _csc(A)
        self.am, self.an = A.shape
        b = _f64(b)
        q = _f64(q)
        t, ln = _cones(cones)
        self._h = self._L.fosor_create_feasibility(self.am, self.an, _i(colptr), _i(rowval), _d(nzval), 0, _d(b),
                                                 _d(q), int(beta), 1 if decreasing_accuracy else 0, len(t),
                                                 _i32(t), _i(ln))
        if not self._h:
            raise ValueError("cones do not cover 1:(an+am)")
        self.N = self._L.fosor_iterate_length(self._h)

    def initial_value(self):
        return np.zeros(self.N)  # Feasibility.jl:57-58


def cg_csc(A, b, x0, tol=None, max_iters=10000):
    """conjugategradient!(x, A, b, ...) on a sparse/dense matrix (test/conjugateGradient.jl)."""
    A, colptr, rowval, nzval = _csc(A)
    x = _f64(x0).copy()
    b = _f64(b)
    it = lib().fosor_cg_csc(A.shape[0], _i(colptr), _i(rowval), _d(nzval), 0, _d(b), _d(x),
                            -1.0 if tol is None else float(tol), int(max_iters))
    return x, int(it)


def prox_cone(name, x, dual=False):
    x = _f64(x)
    y = np.empty_like(x)
    rc = lib().fosor_prox_cone(CONE_CODES[name], 1 if dual else 0, _d(x), _d(y), x.size)
    if rc != 0:
        raise NotImplementedError(f"cone {name} not restated in the oracle")
    return y


def host_threads(variant: str = "") -> int:
    """1 for the restatement proper (serial, like the reference's mat-vecs and broadcasts); the OpenMP thread
    count for the "mt" timing variant."""
    return int(lib(variant).fosor_threads()) if variant == "mt" else 1


class OracleConicDenseBig(_OracleBase):
    """bench.py's config 2 at FULL size on the host: a dense m x n matrix stored the way the reference stores it
    (SparseMatrixCSC{Float64,Int64}, every entry present: 16 B per entry), generated in place by the C library
    (counter-based N(0,1)/sqrt(n)), never copied.  b = A xi + s, c = -A' y with xi, s, y from ``vectors(m, n)``."""

    def __init__(self, m, n, seed, vectors, variant="mt"):
        super().__init__(variant)
        L = self._L
        self.m, self.n = int(m), int(n)
        self.colptr = np.empty(n + 1, dtype=np.int64)
        self.rowval = np.empty(m * n, dtype=np.int64)
        self.nzval = np.empty(m * n, dtype=np.float64)
        L.fosor_gen_dense_csc(m, n, int(seed), 1.0 / np.sqrt(n), _i(self.colptr), _i(self.rowval), _d(self.nzval))
        xi, s, y, cones = vectors(m, n, seed)
        b = np.empty(m)
        c = np.empty(n)
        L.fosor_csc_mul(m, n, _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(_f64(xi)), _d(b), 0)
        L.fosor_csc_mul(m, n, _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(_f64(y)), _d(c), 1)
        self.b = b + s
        self.c = -c
        t1, l1 = _cones(cones)
        t2, l2 = _cones([("Free", n)])
        self._h = L.fosor_create_conic_borrowed(m, n, _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(self.b),
                                                _d(self.c), len(t1), _i32(t1), _i(l1), len(t2), _i32(t2), _i(l2))
        if not self._h:
            raise ValueError("cones do not cover 1:m / 1:n")
        self.N = L.fosor_iterate_length(self._h)
        self.l = m + n + 1

    def initial_value(self):
        z = np.zeros(self.N)
        z[self.l - 1] = 1.0
        z[2 * self.l - 1] = 1.0
        return z


__all__ = ["OracleConic", "OracleFeasibility", "OracleConicDenseBig", "build_all", "VARIANTS", "cg_csc", "prox_cone", "build", "lib", "records_to_history",
           "CONE_CODES", "ALG_CODES", "STATUS_NAMES", "REC_FIELDS", "host_threads"]

if __name__ == "__main__":
    print(build(force=bool(os.environ.get("FORCE"))))
