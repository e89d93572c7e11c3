"""Host-side mirror of the reference's model / driver layer, running its hot loop on the B200
through the C ABI (``include/fos_b200.h``).

Mirrors (paths relative to /root/reference/src):

* ``FOSMathProgModel``, ``Solution``                      types.jl:6-60
* ``ConicModel``, ``loadproblem!``, ``optimize!``, ``status``, ``getobjval``, ``getsolution``,
  ``numvar``, ``numconstr``, ``supportedcones``           FOSSolverInterface.jl:5-69
* ``solve!(model)`` / ``iterate``                          solverwrapper.jl:2-41
* ``HSDEStatus`` printing and history                      problemforms/HSDE/HSDEStatus.jl:73-139
* ``Feasibility``, ``FeasibilityModel``, ``solve!``        problemforms/Feasibility/Feasibility.jl
* ``FeasibilityStatus`` printing and history               problemforms/Feasibility/FeasibilityStatus.jl

Python cannot spell ``loadproblem!``; the bang is dropped, everything else keeps its name,
argument order and error behaviour.  History keys are the reference's (``:p :d :g :ctx :bty :κ
:τ :t :cgiter :x :y :s`` and ``:err :t :z``), stored in an ``MVHistory`` look-alike.
"""
from __future__ import annotations

import ctypes as C
import sys
import time
import warnings
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .algorithms import FOSAlgorithm

CONE_CODES = {"Free": 0, "Zero": 1, "NonNeg": 2, "NonPos": 3, "SOC": 4, "SOCRotated": 5, "SDP": 6,
              "ExpPrimal": 7, "ExpDual": 8}
STATUS_SYMBOLS = {0: "Continue", 1: "Optimal", 2: "Unbounded", 3: "Infeasible", 4: "Indeterminate"}
REC_LEN = _lib.FOS_REC_LEN

_dp = _lib._dp


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))


class MVHistory:
    """Minimal ValueHistories.MVHistory: ``push!(h, key, i, v)`` / ``get(h, key) -> (is, vs)``."""

    def __init__(self):
        self._d = {}

    def push(self, key, i, v):
        it, vs = self._d.setdefault(key, ([], []))
        it.append(i)
        vs.append(v)

    def get(self, key):
        it, vs = self._d[key]
        return list(it), list(vs)

    def keys(self):
        return list(self._d.keys())

    def __contains__(self, key):
        return key in self._d

    def __len__(self):
        return len(self._d)


def _cone_arrays(cones, total, what):
    """[(symbol, indices-or-length), ...] -> (types, lens).  Index vectors must be contiguous,
    ordered and cover 1:total (cones.jl:44-56, 66-72)."""
    types, lens = [], []
    prev_end = 0
    for sym, idx in cones:
        name = sym.lstrip(":") if isinstance(sym, str) else sym
        if name not in CONE_CODES:
            raise KeyError(f"unknown cone {sym!r}")
        if isinstance(idx, (int, np.integer)):
            start, ln = prev_end + 1, int(idx)
        else:
            arr = np.asarray(list(idx) if isinstance(idx, range) else idx, dtype=np.int64).reshape(-1)
            if arr.size == 0:
                start, ln = prev_end + 1, 0
            else:
                if not np.array_equal(arr, np.arange(arr[0], arr[-1] + 1)):
                    raise ValueError("Invalid range in input")  # cones.jl:50
                start, ln = int(arr[0]), int(arr.size)
        if start != prev_end + 1:
            raise AssertionError(f"{what} cones must be contiguous and ordered (cones.jl:69)")
        prev_end = start + ln - 1
        types.append(CONE_CODES[name])
        lens.append(ln)
    if prev_end != total:
        raise AssertionError(f"{what} cones must cover 1:{total} (cones.jl:66-72)")
    return np.array(types, dtype=np.int32), np.array(lens, dtype=np.int64)


def _i32p(a):
    return a.ctypes.data_as(_lib._i32p)


def _i64p(a):
    return a.ctypes.data_as(_lib._i64p)


class _Handle:
    """RAII wrapper of a ``fos_handle_t``."""

    def __init__(self, device=0):
        self.L = _lib.load()
        self._B = 0
        self.h = C.c_void_p()
        rc = self.L.fos_create(C.byref(self.h), int(device))
        if rc != 0:
            msg = self.L.fos_last_error(None)
            raise _lib.FosError(rc, msg.decode() if msg else "fos_create failed")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.fos_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ck(self, rc):
        _lib.check(self.h, rc)

    # -- thin typed wrappers ------------------------------------------------------------------
    def set_option(self, key, value):
        self.ck(self.L.fos_set_option(self.h, key.encode(), float(value)))

    def n(self):
        n = int(self.L.fos_iterate_length(self.h))
        if n < 0:
            raise _lib.FosError(-1, "no problem loaded on this handle")
        return n

    def set_algorithm(self, alg: FOSAlgorithm):
        code, a, a1, a2, b, ip = alg._params()
        self.ck(self.L.fos_set_algorithm(self.h, code, a, a1, a2, b, ip))
        # LineSearchWrapper(alg; lsinterval) (wrappers/linesearch.jl): only GAP / GAPA take part
        ls = int(getattr(alg, "lsinterval", 0)) if code in (0, 1) else 0
        if self.batch_size() < 0:
            self.ck(self.L.fos_set_linesearch(self.h, ls))

    def set_linesearch(self, lsinterval):
        self.ck(self.L.fos_set_linesearch(self.h, int(lsinterval)))

    def set_direct(self, on=True):
        self.ck(self.L.fos_set_direct(self.h, 1 if on else 0))

    def set_iterate(self, z):
        z = _f64(z)
        self.ck(self.L.fos_set_iterate(self.h, _d(z), z.size))

    def set_initial_iterate(self):
        self.ck(self.L.fos_set_initial_iterate(self.h))

    def get_iterate(self):
        z = np.empty(self.n())
        self.ck(self.L.fos_get_iterate(self.h, _d(z), z.size))
        return z

    _STATE = {"x": 0, "tmp1": 1, "tmp2": 2, "xinit": 3, "rhs": 4, "proj": 5, "fista_y": 6, "dykstra_p": 7,
              "dykstra_q": 8}
    _INFO = {"s1_calls": 0, "cgiter": 1, "alpha12": 2, "fista_t": 3, "cg_warned": 4, "total_cg": 5,
             "total_passes": 6, "launches": 7, "alphabest": 8, "mv2_ms": 9, "mv2_n": 10, "mv1_ms": 11, "mv1_n": 12,
             "mv_skipped": 13, "bytes_per_pass": 14, "num_sms": 15, "tail_ms": 16, "tail_n": 17,
             "storage_kind": 18, "hybrid_dense_rows": 19, "hybrid_sparse_rows": 20, "k1_balanced": 21,
             "k1_spread_before": 22, "k1_spread_after": 23}

    def get_state(self, which):
        z = np.empty(self.n())
        self.ck(self.L.fos_get_state(self.h, self._STATE[which], _d(z), z.size))
        return z

    def set_state(self, which, z):
        z = _f64(z)
        self.ck(self.L.fos_set_state(self.h, self._STATE[which], _d(z), z.size))

    def set_info(self, which, value):
        self.ck(self.L.fos_set_info(self.h, self._INFO[which], float(value)))

    def info(self, which):
        out = C.c_double(0)
        self.ck(self.L.fos_get_info(self.h, self._INFO[which], C.byref(out)))
        return out.value

    def run(self, i_start, n_iters, checki, eps, trace=False):
        cap = n_iters // max(checki, 1) + 2
        rec = np.zeros((cap, REC_LEN))
        done = C.c_int64(0)
        st = C.c_int32(0)
        nrec = C.c_int64(0)
        tr = np.zeros((n_iters, self.n())) if trace else None
        self.ck(self.L.fos_run(self.h, i_start, n_iters, checki, eps, C.byref(done), C.byref(st), _d(rec), cap,
                               C.byref(nrec), _d(tr) if trace else None))
        return int(done.value), int(st.value), rec[:min(nrec.value, cap)], (tr[:done.value] if trace else None)

    def finish(self):
        guess = np.empty(self.n())
        rec = np.zeros((1, REC_LEN))
        nrec = C.c_int64(0)
        st = C.c_int32(0)
        self.ck(self.L.fos_finish(self.h, _d(guess), guess.size, _d(rec), C.byref(nrec), C.byref(st)))
        return guess, rec[:nrec.value], int(st.value)

    def solve(self, max_iters, checki, eps):
        cap = max_iters // max(checki, 1) + 2
        rec = np.zeros((cap, REC_LEN))
        guess = np.empty(self.n())
        done = C.c_int64(0)
        st = C.c_int32(0)
        nrec = C.c_int64(0)
        self.ck(self.L.fos_solve(self.h, max_iters, checki, eps, _d(guess), guess.size, C.byref(done), C.byref(st),
                                 _d(rec), cap, C.byref(nrec)))
        return int(done.value), int(st.value), rec[:min(nrec.value, cap)], guess

    # -- batch mode (fos_*_batch): every array has a leading batch dimension -----------------------
    def load_conic_batch(self, A, b, c, constr_cones, var_cones, device_ptr=None):
        """A: (B, m, n) ndarray (or its shape when ``device_ptr=(ptr, lda, pstride)`` hands over
        matrices already on the GPU); b: (B, m); c: (B, n); cones shared by all problems."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        B, m = b.shape
        n = c.shape[1]
        t1, l1 = _cone_arrays(constr_cones, m, "constraint")
        t2, l2 = _cone_arrays(var_cones, n, "variable")
        if device_ptr is not None:
            ptr, lda, pstride = device_ptr
            loc = 1
            keep = None
        else:
            keep = np.ascontiguousarray(A, dtype=np.float64)
            assert keep.shape == (B, m, n), (keep.shape, (B, m, n))
            ptr, lda, pstride, loc = keep.ctypes.data, n, m * n, 0
        self.ck(self.L.fos_load_conic_dense_batch(self.h, B, m, n, C.c_void_p(int(ptr)), int(lda), int(pstride), loc,
                                                  _d(b), _d(c), len(t1), _i32p(t1), _i64p(l1), len(t2), _i32p(t2),
                                                  _i64p(l2)))
        self._B = B

    def batch_size(self):
        return int(self.L.fos_batch_size(self.h))

    def set_iterate_batch(self, z=None):
        if z is None:
            self.ck(self.L.fos_set_iterate_batch(self.h, None))
        else:
            z = np.ascontiguousarray(z, dtype=np.float64)
            assert z.shape == (self._B, self.n())
            self.ck(self.L.fos_set_iterate_batch(self.h, _d(z)))

    def get_iterate_batch(self):
        z = np.empty((self._B, self.n()))
        self.ck(self.L.fos_get_iterate_batch(self.h, _d(z)))
        return z

    def get_state_batch(self, which):
        z = np.empty((self._B, self.n()))
        self.ck(self.L.fos_get_state_batch(self.h, self._STATE[which], _d(z)))
        return z

    def set_state_batch(self, which, z):
        z = np.ascontiguousarray(z, dtype=np.float64)
        assert z.shape == (self._B, self.n())
        self.ck(self.L.fos_set_state_batch(self.h, self._STATE[which], _d(z)))

    def info_batch(self, which):
        out = np.empty(self._B)
        self.ck(self.L.fos_get_info_batch(self.h, self._INFO[which], _d(out)))
        return out

    def set_info_batch(self, which, values):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), (self._B,)))
        self.ck(self.L.fos_set_info_batch(self.h, self._INFO[which], _d(v)))

    def run_batch(self, i_start, n_iters, checki, eps):
        B = self._B
        cap = n_iters // max(checki, 1) + 2
        rec = np.zeros((B, cap, REC_LEN))
        done = np.zeros(B, dtype=np.int64)
        st = np.zeros(B, dtype=np.int32)
        nrec = np.zeros(B, dtype=np.int64)
        self.ck(self.L.fos_run_batch(self.h, i_start, n_iters, checki, eps, _i64p(done), _i32p(st), _d(rec), cap,
                                     _i64p(nrec)))
        return done, st, [rec[p, :min(nrec[p], cap)] for p in range(B)]

    def finish_batch(self):
        B = self._B
        guess = np.empty((B, self.n()))
        rec = np.zeros((B, REC_LEN))
        nrec = np.zeros(B, dtype=np.int64)
        st = np.zeros(B, dtype=np.int32)
        self.ck(self.L.fos_finish_batch(self.h, _d(guess), _d(rec), _i64p(nrec), _i32p(st)))
        return guess, [rec[p:p + 1][:nrec[p]] for p in range(B)], st

    def solve_batch(self, max_iters, checki, eps):
        B = self._B
        cap = max_iters // max(checki, 1) + 2
        rec = np.zeros((B, cap, REC_LEN))
        guess = np.empty((B, self.n()))
        done = np.zeros(B, dtype=np.int64)
        st = np.zeros(B, dtype=np.int32)
        nrec = np.zeros(B, dtype=np.int64)
        self.ck(self.L.fos_solve_batch(self.h, max_iters, checki, eps, _d(guess), _i64p(done), _i32p(st), _d(rec), cap,
                                       _i64p(nrec)))
        return done, st, [rec[p, :min(nrec[p], cap)] for p in range(B)], guess

    # unit level
    def _vec_call(self, fn, x, n_out=None, *extra):
        x = _f64(x)
        y = np.empty(self.n() if n_out is None else n_out)
        self.ck(fn(self.h, _d(x), _d(y), *extra))
        return y

    def kkt_mul(self, x):
        return self._vec_call(self.L.fos_kkt_mul, x)

    def affine_prox(self, x):
        return self._vec_call(self.L.fos_affine_prox, x)

    def hsdematrix_prox(self, x):
        return self._vec_call(self.L.fos_hsdematrix_prox, x)

    def cone_prox(self, x):
        return self._vec_call(self.L.fos_cone_prox, x)

    def q_mul(self, B, transpose=False):
        return self._vec_call(self.L.fos_q_mul, B, self.n() // 2, 1 if transpose else 0)

    def a_mul(self, x, m, n, transpose=False):
        return self._vec_call(self.L.fos_a_mul, x, n if transpose else m, 1 if transpose else 0)

    def cg_dense(self, A, b, x0, tol=None, max_iters=10000):
        A = np.ascontiguousarray(A, dtype=np.float64)
        b = _f64(b)
        x = _f64(x0).copy()
        it = C.c_int64(0)
        self.ck(self.L.fos_cg_dense(self.h, A.shape[0], _d(A), _d(b), _d(x), -1.0 if tol is None else float(tol),
                                    int(max_iters), C.byref(it)))
        return x, int(it.value)

    def prox_cone(self, name, x, dual=False):
        x = _f64(x)
        y = np.empty_like(x)
        self.ck(self.L.fos_prox_cone(self.h, CONE_CODES[name], 1 if dual else 0, _d(x), _d(y), x.size))
        return y

    def stream(self):
        out = C.c_uint64(0)
        self.ck(self.L.fos_get_stream(self.h, C.byref(out)))
        return int(out.value)

    def time_psd(self, X, reps=5):
        """X: (ncones, d(d+1)/2) packed matrices -> (projections, ms per call, Jacobi sweeps)."""
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        nc, plen = X.shape
        d = int(round(np.sqrt(0.25 + 2 * plen) - 0.5))
        Y = np.empty_like(X)
        ms = C.c_double(0)
        sw = C.c_int32(0)
        self.ck(self.L.fos_time_psd(self.h, d, nc, _d(X), _d(Y), int(reps), C.byref(ms), C.byref(sw)))
        return Y, ms.value, int(sw.value)

    def tail_trace(self):
        """Phase timing of the fused CG tail (option "tail_trace" = 1): (3 blocks, 16) summed SM cycles;
        column 15 = launches counted (fos_get_tail_trace)."""
        out = np.zeros(48)
        self.ck(self.L.fos_get_tail_trace(self.h, _d(out)))
        return out.reshape(3, 16)

    def time_matvec(self, nvec=2, reps=10):
        ms = C.c_double(0)
        by = C.c_double(0)
        self.ck(self.L.fos_time_matvec(self.h, nvec, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value


# =============================================================================================
# conic model (MathProgBase path)
# =============================================================================================
@dataclass
class Solution:  # types.jl:6-11
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    status: str


def _csc_arrays(A):
    import scipy.sparse as sp
    A = sp.csc_matrix(A, dtype=np.float64)
    A.sort_indices()
    return (A, np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int64),
            np.ascontiguousarray(A.data, dtype=np.float64))


class FOSMathProgModel:
    """types.jl:30-60.  ``data`` is the device handle that replaces GAPData/GAPAData/..."""

    def __init__(self, s: FOSAlgorithm, **kwargs):
        self.input_numconstr = 0
        self.input_numvar = 0
        self.K1 = ()
        self.K2 = ()
        self.A = None
        self.b = np.zeros(0)
        self.c = np.zeros(0)
        self.alg = s
        self.data = None
        self.solve_stat = "NotSolved"
        self.obj_val = 0.0
        self.primal_sol = np.zeros(0)
        self.dual_sol = np.zeros(0)
        self.slack = np.zeros(0)
        self.options = dict(kwargs)
        self.enditr = -1  # never written by the reference either (SURVEY a-Q 12)
        self.init_duration = 1  # ns placeholder, types.jl:59
        self.history = MVHistory()
        self.device = int(self.options.get("device", 0))
        self.storage = self.options.get("storage", "auto")


def ConicModel(s: FOSAlgorithm) -> FOSMathProgModel:
    """FOSSolverInterface.jl:5"""
    return FOSMathProgModel(s, **s.options)


def supportedcones(s: FOSAlgorithm):
    """FOSSolverInterface.jl:69 (``:SOCRotated`` is missing there as well, SURVEY a-Q 9)."""
    return ["Free", "Zero", "NonNeg", "NonPos", "SOC", "SDP", "ExpPrimal", "ExpDual"]


def loadproblem(model: FOSMathProgModel, c, A, b, constr_cones, var_cones, *, device_matrix=None):
    """``loadproblem!`` (FOSSolverInterface.jl:27-64).  ``A`` may be dense or any SciPy sparse
    matrix; it is lowered to CSC like the reference's ``sparse(A)`` unless the model was created
    with ``storage="dense"`` or ``A`` is a dense ndarray above 25 % fill.

    ``device_matrix=(ptr, lda)`` hands over a row-major FP64 matrix that already lives on the GPU
    (benchmark shapes); ``A`` is then only used for its ``shape``.
    """
    t1 = time.perf_counter_ns()
    model.alg._check_supported()
    c = _f64(c)
    b = _f64(b)
    m, n = (A.shape if hasattr(A, "shape") else np.asarray(A).shape)
    model.input_numconstr, model.input_numvar = m, n
    t1a, l1a = _cone_arrays(constr_cones, m, "constraint")
    t2a, l2a = _cone_arrays(var_cones, n, "variable")
    model.K1 = tuple(constr_cones)
    model.K2 = tuple(var_cones)
    model.A, model.b, model.c = A, b, c
    H = _Handle(model.device)
    for k, v in model.options.items():
        if k in ("matvec_impl", "grid_ctas", "cg_batch"):
            H.set_option(k, v)
    storage = {"auto": 0, "dense": 1, "sparse": 2}[model.storage]
    if device_matrix is not None:
        ptr, lda = device_matrix
        H.ck(H.L.fos_load_conic_dense(H.h, m, n, C.c_void_p(int(ptr)), int(lda), 1, 0, m, _d(b), _d(c), len(t1a),
                                      _i32p(t1a), _i64p(l1a), len(t2a), _i32p(t2a), _i64p(l2a)))
    elif isinstance(A, np.ndarray) and storage != 2 and (storage == 1 or np.count_nonzero(A) > 0.25 * A.size):
        Ad = np.ascontiguousarray(A, dtype=np.float64)
        H.ck(H.L.fos_load_conic_dense(H.h, m, n, Ad.ctypes.data_as(C.c_void_p), n, 0, 0, m, _d(b), _d(c), len(t1a),
                                      _i32p(t1a), _i64p(l1a), len(t2a), _i32p(t2a), _i64p(l2a)))
    else:
        _, colptr, rowval, nzval = _csc_arrays(A)
        H.ck(H.L.fos_load_conic_csc(H.h, m, n, _i64p(colptr), _i64p(rowval), _d(nzval), 0, _d(b), _d(c), len(t1a),
                                    _i32p(t1a), _i64p(l1a), len(t2a), _i32p(t2a), _i64p(l2a), storage))
    if model.alg.direct:  # HSDE(model, direct=alg.direct)  FOSSolverInterface.jl:77
        H.set_direct(True)
    H.set_algorithm(model.alg)  # init_algorithm!(model.alg, model)  :58
    model.data = H
    model._form = "hsde"
    model.init_duration = time.perf_counter_ns() - t1
    return model


def numvar(model):
    return model.input_numvar


def numconstr(model):
    return model.input_numconstr


def status(model):
    return model.solve_stat


def getobjval(model):
    return model.obj_val


def getsolution(model):
    return model.primal_sol.copy()


# ---- printing (HSDEStatus.jl:73-91, FeasibilityStatus.jl:74-92) ---------------------------------
def _jl_float(x):
    r = repr(float(x))
    if "e" in r:
        mant, ex = r.split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}e{int(ex)}"
    return r


def _print_header_hsde(init_duration_ns, direct, out):
    print(f"Time to initialize: {_jl_float(init_duration_ns / 1e9)}s", file=out)
    width = 76 + (0 if direct else 5)
    print("-" * width, file=out)
    line = " Iter | pri res | dua res | rel gap | pri obj | dua obj | kap/tau"
    if not direct:
        line += " | cg "
    print(line + " | time", file=out)
    print("-" * width, file=out)


def _print_iter_hsde(i, p, d, g, ctx, bty, kt, cgiter, t, direct, out):
    if direct:
        print("%6d|% 9.2e % 9.2e % 9.2e % 9.2e % 9.2e % 9.2e % .1es" % (i, p, d, g, ctx, -bty, kt, t / 1e9), file=out)
    else:
        print("%6d|% 9.2e % 9.2e % 9.2e % 9.2e % 9.2e % 9.2e % 4d % .1es" % (i, p, d, g, ctx, -bty, kt, cgiter,
                                                                            t / 1e9), file=out)


def _print_header_feas(init_duration_ns, direct, out):
    print(f"Time to initialize: {_jl_float(init_duration_ns / 1e9)}s", file=out)
    width = 22 + (0 if direct else 5)
    print("-" * width, file=out)
    line = " Iter | res"
    if not direct:
        line += " | cg "
    print(line + " | time", file=out)
    print("-" * width, file=out)


def _solve_model(model, out=None):
    """solve!(model) + iterate (solverwrapper.jl:2-41): the loop runs on the device, up to `checki`
    iterations per library call; history, printing and the warning surface here."""
    out = out or sys.stdout
    opts = dict(model.options)
    max_iters = int(opts.get("max_iters", 10000))  # solverwrapper.jl:5-10
    verbose = int(opts.get("verbose", 1))
    debug = int(opts.get("debug", 1))
    eps = float(opts.get("eps", 1e-5))
    checki = int(opts.get("checki", 100))
    H: _Handle = model.data
    hsde = model._form == "hsde"
    if "initx" in opts:
        H.set_iterate(opts["initx"])
    else:
        H.set_initial_iterate()
    H.ck(H.L.fos_begin_solve(H.h))  # status = model.status_generator(...)  solverwrapper.jl:13
    init_time = time.perf_counter_ns()
    t1 = time.time()
    if verbose > 0:  # printstatusheader
        if hsde:
            # the HSDE closure captured the placeholder init_duration = 1 ns (SURVEY a-Q 11)
            _print_header_hsde(1, bool(getattr(model.alg, "direct", False)), out)
        else:
            _print_header_feas(model.init_duration, True, out)  # Feasibility.jl:76: direct = true
    i, st, last_i = 1, 0, 0
    while i <= max_iters and st == 0:
        chunk = min(checki - ((i - 1) % checki), max_iters - i + 1)
        done, st, rec, _ = H.run(i, chunk, checki, eps)
        last_i = i + done - 1
        i += done
        t = time.perf_counter_ns() - init_time
        for r in rec:
            model._record(r, t, verbose, debug, out)
        if done < chunk:
            break
    guess, rec, st = H.finish()  # getsol + forced final check, solverwrapper.jl:31-34
    t = time.perf_counter_ns() - init_time
    for r in rec:
        model._record(r, t, verbose, debug, out)
    if H.info("cg_warned"):
        warnings.warn("CG reached max iterations, result may be inaccurate")  # conjugategradients.jl:53
    if verbose > 0:
        print("Time for iterations: ", file=out)
        print(f"{time.time() - t1} s", file=out)
    model.last_iteration = last_i
    return guess, st


def _hsde_record(model, r, t, verbose, debug, out):
    i = int(r[0])
    p, d, g, ctx, bty, kap, tau, cgiter, st = r[1], r[2], r[3], r[4], r[5], r[6], r[7], int(r[8]), int(r[9])
    h = model.history
    if debug > 0:  # savedata, HSDEStatus.jl:125-139
        for key, v in (("p", p), ("d", d), ("g", g), ("ctx", ctx), ("bty", bty), ("κ", kap), ("τ", tau), ("t", t)):
            h.push(key, i, v)
        if debug > 1:
            # copies, where the reference stores aliasing views (SURVEY a-Q 4)
            z = model.data.get_state("proj")
            n, m = model.input_numvar, model.input_numconstr
            l = n + m + 1
            h.push("x", i, z[:n].copy())
            h.push("y", i, z[n:n + m].copy())
            h.push("s", i, z[l + n:l + n + m].copy())
    if verbose > 0:
        direct = bool(getattr(model.alg, "direct", False))
        if not direct:
            h.push("cgiter", i, cgiter)  # HSDEStatus.jl:43-47
        with np.errstate(all="ignore"):
            kt = np.float64(kap) / np.float64(tau)
        _print_iter_hsde(i, p, d, g, ctx, bty, kt, cgiter, t, direct, out)
        if st == 1:
            print(f"Found solution i={i}", file=out)  # HSDEStatus.jl:56


FOSMathProgModel._record = _hsde_record


def optimize(model: FOSMathProgModel, out=None):
    """``optimize!`` (FOSSolverInterface.jl:8-21)."""
    model.history = MVHistory()  # :10
    guess, st = _solve_model(model, out)
    # HSDE_populatesolution (HSDE.jl:49-61)
    n, m = model.input_numvar, model.input_numconstr
    l = n + m + 1
    tau = guess[l - 1]
    endstatus = STATUS_SYMBOLS[st]
    if endstatus == "Continue":
        endstatus = "Indeterminate"
    with np.errstate(all="ignore"):
        sol = Solution(guess[:n] / tau, guess[n:n + m] / tau, guess[l + n:l + n + m] / tau, endstatus)
    model.solve_stat = sol.status
    model.primal_sol = sol.x
    model.dual_sol = sol.y
    model.slack = sol.s
    model.obj_val = float(np.dot(model.c, model.primal_sol))  # :20
    return model


# =============================================================================================
# batch of conic models (config 5: many independent problems of one shape)
# =============================================================================================
def solve_batch(alg: FOSAlgorithm, cs, As, bs, constr_cones, var_cones, device=0, device_ptr=None):
    """``[solve!(model_j) for j in 1:B]`` for B models that share (m, n, cones) -- loadproblem! +
    optimize! (FOSSolverInterface.jl:8-64) with the whole batch resident on one GPU and one persistent
    CTA per problem.  Options (max_iters, eps, checki, debug) come from ``alg.options`` as for
    ``ConicModel``.  Returns a list of ``FOSMathProgModel`` with solve_stat / primal_sol / dual_sol /
    slack / obj_val / history filled in, in input order."""
    if alg.direct:
        raise NotImplementedError("direct=true is not offered in batch mode; construct the algorithm with direct=False")
    opts = dict(alg.options)
    max_iters = int(opts.get("max_iters", 10000))
    eps = float(opts.get("eps", 1e-5))
    checki = int(opts.get("checki", 100))
    debug = int(opts.get("debug", 1))
    verbose = int(opts.get("verbose", 1))
    bs = np.ascontiguousarray(bs, dtype=np.float64)
    cs = np.ascontiguousarray(cs, dtype=np.float64)
    B, m = bs.shape
    n = cs.shape[1]
    H = _Handle(device)
    if "batch_ctas" in opts:
        H.set_option("batch_ctas", opts["batch_ctas"])
    H.load_conic_batch(As, bs, cs, constr_cones, var_cones, device_ptr=device_ptr)
    H.set_algorithm(alg)
    if "initx" in opts:
        H.set_iterate_batch(np.broadcast_to(np.asarray(opts["initx"], dtype=np.float64), (B, H.n())))
    t0 = time.perf_counter_ns()
    done, st, recs, guess = H.solve_batch(max_iters, checki, eps)
    t = time.perf_counter_ns() - t0
    if H.info("cg_warned"):
        warnings.warn("CG reached max iterations, result may be inaccurate")  # conjugategradients.jl:53
    models = []
    l = n + m + 1
    for j in range(B):
        mod = FOSMathProgModel(alg, **alg.options)
        mod.input_numconstr, mod.input_numvar = m, n
        mod.K1, mod.K2 = tuple(constr_cones), tuple(var_cones)
        mod.b, mod.c = bs[j], cs[j]
        mod.last_iteration = int(done[j])
        for r in recs[j]:
            if debug > 0:   # savedata (HSDEStatus.jl:125-139); :t is the wall time of the WHOLE batch launch: the
                # problems are solved concurrently, so there is no per-check clock to report
                for key, v in (("p", r[1]), ("d", r[2]), ("g", r[3]), ("ctx", r[4]), ("bty", r[5]), ("κ", r[6]),
                               ("τ", r[7]), ("t", t)):
                    mod.history.push(key, int(r[0]), v)
            if verbose > 0:  # HSDEStatus.jl:43-47: cgiter is gated on verbose, not on debug
                mod.history.push("cgiter", int(r[0]), int(r[8]))
        g = guess[j]
        tau = g[l - 1]
        with np.errstate(all="ignore"):
            mod.primal_sol, mod.dual_sol, mod.slack = g[:n] / tau, g[n:n + m] / tau, g[l + n:l + n + m] / tau
        stj = int(st[j])
        mod.solve_stat = "Indeterminate" if STATUS_SYMBOLS[stj] == "Continue" else STATUS_SYMBOLS[stj]  # HSDE.jl:56-59
        mod.obj_val = float(np.dot(cs[j], mod.primal_sol))
        models.append(mod)
    return models


# =============================================================================================
# Feasibility form
# =============================================================================================
@dataclass
class AffinePlusLinear:
    """``AffinePlusLinear(A, b, q, β; decreasing_accuracy=false)`` utilities/affinepluslinear.jl:71-79:
    f([x;z]) = q'x + i(Ax - βz == b)."""
    A: object
    b: np.ndarray
    q: np.ndarray
    β: int
    decreasing_accuracy: bool = False


@dataclass
class ConeProduct:
    """``ConeProduct(ranges, cones)`` cones.jl:31-77 as [(symbol, range-or-length), ...]."""
    cones: list


@dataclass
class Feasibility:  # Feasibility.jl:2-6
    S1: AffinePlusLinear
    S2: ConeProduct
    n: int


@dataclass
class FeasibilitySolution:  # Feasibility.jl:8-11
    x: np.ndarray
    status: str


class FeasibilityModel:  # Feasibility.jl:15-49
    def __init__(self, problem: Feasibility, alg: FOSAlgorithm, **kwargs):
        t1 = time.perf_counter_ns()
        alg._check_supported_feas()
        self.S1, self.S2, self.n = problem.S1, problem.S2, problem.n
        self.alg = alg
        self.solve_stat = "NotSolved"
        self.obj_val = 0.0
        allkw = dict(alg.options)
        allkw.update(kwargs)  # call-site kwargs override alg.options, Feasibility.jl:33-36
        self.options = allkw
        self.enditr = -1
        self.history = MVHistory()
        S1 = problem.S1
        if not isinstance(S1, AffinePlusLinear) or not isinstance(problem.S2, ConeProduct):
            raise NotImplementedError("the B200 path needs S1::AffinePlusLinear and S2::ConeProduct; arbitrary "
                                      "ProximableFunctions cannot cross the C ABI (SURVEY.md section 2)")
        Am, colptr, rowval, nzval = _csc_arrays(S1.A)
        am, an = Am.shape
        if am + an != problem.n:
            raise AssertionError("Feasibility.n must equal an + am")
        # ("Box", range-or-length, lo, hi) entries are IndBox(lo, hi) (ProximalOperators): loaded as Free, then
        # turned into clamps with fos_set_box
        plain, boxes, pos = [], [], 0
        for entry in problem.S2.cones:
            name = entry[0].lstrip(":") if isinstance(entry[0], str) else entry[0]
            cnt = int(entry[1]) if isinstance(entry[1], (int, np.integer)) else len(entry[1])
            if name == "Box":
                boxes.append((pos, cnt, float(entry[2]), float(entry[3])))
                plain.append(("Free", entry[1]))
            else:
                plain.append((entry[0], entry[1]))
            pos += cnt
        t, ln = _cone_arrays(plain, am + an, "S2")
        b = _f64(S1.b)
        q = _f64(S1.q)
        H = _Handle(int(allkw.get("device", 0)))
        storage = {"auto": 0, "dense": 1, "sparse": 2}[allkw.get("storage", "auto")]
        H.ck(H.L.fos_load_affine_csc(H.h, am, an, _i64p(colptr), _i64p(rowval), _d(nzval), 0, _d(b), _d(q),
                                     int(S1.β), 1 if S1.decreasing_accuracy else 0, len(t), _i32p(t), _i64p(ln),
                                     storage))
        for (start, cnt, lo, hi) in boxes:
            H.ck(H.L.fos_set_box(H.h, start, cnt, lo, hi))
        H.set_algorithm(alg)
        self.data = H
        self._form = "feas"
        self.init_duration = time.perf_counter_ns() - t1

    def _record(self, r, t, verbose, debug, out):
        i, err, cgiter, st = int(r[0]), r[1], int(r[8]), int(r[9])
        h = self.history
        if debug > 0:  # FeasibilityStatus.jl:94-103
            h.push("err", i, err)
            h.push("t", i, t)
            if debug > 1:
                h.push("z", i, self.data.get_state("proj"))
        if verbose > 0:
            print("%6d|% 9.2e % .1es" % (i, err, t / 1e9), file=out)  # direct = true printing, Feasibility.jl:76
            if st == 1:
                print(f"Found solution i={i}", file=out)


def _check_supported_feas(self):
    # the Feasibility form takes its S1 from the user (Feasibility.jl:75-81): alg.direct is ignored there
    return None


FOSAlgorithm._check_supported_feas = _check_supported_feas


def solve(problem: Feasibility, alg: FOSAlgorithm, out=None, **kwargs):
    """``solve!(problem::Feasibility, alg; kwargs...) -> (solution, model)`` Feasibility.jl:51-55."""
    model = FeasibilityModel(problem, alg, **kwargs)
    guess, st = _solve_model(model, out)
    endstatus = STATUS_SYMBOLS[st]
    if endstatus == "Continue":
        endstatus = "Indeterminate"
    model.solve_stat = endstatus
    return FeasibilitySolution(guess, endstatus), model
