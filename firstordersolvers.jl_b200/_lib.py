"""ctypes binding of ``libfos_b200.so`` -- exactly the entry points of ``include/fos_b200.h``.

There is no CPU fallback: if the shared library is missing it is built with nvcc (which needs
no GPU); if it cannot be built or loaded, or if no sm_100 device is present when a handle is
created, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_h = C.c_void_p

FOS_REC_LEN = 10
FOS_COMM_ID_BYTES = 128
FOS_IPC_HANDLE_BYTES = 64

# name -> (restype, argtypes); mirrors include/fos_b200.h one to one
SIGNATURES = {
    "fos_abi_version": (C.c_int32, []),
    "fos_create": (C.c_int32, [C.POINTER(_h), C.c_int32]),
    "fos_destroy": (C.c_int32, [_h]),
    "fos_last_error": (C.c_char_p, [_h]),
    "fos_set_option": (C.c_int32, [_h, C.c_char_p, C.c_double]),
    "fos_comm_unique_id": (C.c_int32, [_u8p]),
    "fos_comm_init": (C.c_int32, [_h, C.c_int32, C.c_int32, _u8p]),
    "fos_comm_p2p_export": (C.c_int32, [_h, _u8p]),
    "fos_comm_p2p_import": (C.c_int32, [_h, _u8p]),
    "fos_load_conic_csc": (C.c_int32, [_h, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int64, _dp, _dp,
                                       C.c_int64, _i32p, _i64p, C.c_int64, _i32p, _i64p, C.c_int32]),
    "fos_load_conic_dense": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                         C.c_int64, _dp, _dp, C.c_int64, _i32p, _i64p, C.c_int64, _i32p, _i64p]),
    "fos_load_affine_csc": (C.c_int32, [_h, C.c_int64, C.c_int64, _i64p, _i64p, _dp, C.c_int64, _dp, _dp,
                                        C.c_int32, C.c_int32, C.c_int64, _i32p, _i64p, C.c_int32]),
    "fos_set_algorithm": (C.c_int32, [_h, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64]),
    "fos_set_box": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_double, C.c_double]),
    "fos_set_linesearch": (C.c_int32, [_h, C.c_int64]),
    "fos_set_direct": (C.c_int32, [_h, C.c_int32]),
    "fos_iterate_length": (C.c_int64, [_h]),
    "fos_set_iterate": (C.c_int32, [_h, _dp, C.c_int64]),
    "fos_set_initial_iterate": (C.c_int32, [_h]),
    "fos_get_iterate": (C.c_int32, [_h, _dp, C.c_int64]),
    "fos_get_state": (C.c_int32, [_h, C.c_int32, _dp, C.c_int64]),
    "fos_set_state": (C.c_int32, [_h, C.c_int32, _dp, C.c_int64]),
    "fos_get_info": (C.c_int32, [_h, C.c_int32, _dp]),
    "fos_set_info": (C.c_int32, [_h, C.c_int32, C.c_double]),
    "fos_begin_solve": (C.c_int32, [_h]),
    "fos_run": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_int64, C.c_double, _i64p, _i32p, _dp, C.c_int64, _i64p,
                            _dp]),
    "fos_finish": (C.c_int32, [_h, _dp, C.c_int64, _dp, _i64p, _i32p]),
    "fos_solve": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_double, _dp, C.c_int64, _i64p, _i32p, _dp, C.c_int64,
                              _i64p]),
    "fos_load_conic_dense_batch": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                               C.c_int32, _dp, _dp, C.c_int64, _i32p, _i64p, C.c_int64, _i32p,
                                               _i64p]),
    "fos_batch_size": (C.c_int64, [_h]),
    "fos_set_iterate_batch": (C.c_int32, [_h, _dp]),
    "fos_get_iterate_batch": (C.c_int32, [_h, _dp]),
    "fos_get_state_batch": (C.c_int32, [_h, C.c_int32, _dp]),
    "fos_set_state_batch": (C.c_int32, [_h, C.c_int32, _dp]),
    "fos_get_info_batch": (C.c_int32, [_h, C.c_int32, _dp]),
    "fos_set_info_batch": (C.c_int32, [_h, C.c_int32, _dp]),
    "fos_begin_solve_batch": (C.c_int32, [_h]),
    "fos_run_batch": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_int64, C.c_double, _i64p, _i32p, _dp, C.c_int64,
                                  _i64p]),
    "fos_finish_batch": (C.c_int32, [_h, _dp, _dp, _i64p, _i32p]),
    "fos_solve_batch": (C.c_int32, [_h, C.c_int64, C.c_int64, C.c_double, _dp, _i64p, _i32p, _dp, C.c_int64, _i64p]),
    "fos_a_mul": (C.c_int32, [_h, _dp, _dp, C.c_int32]),
    "fos_q_mul": (C.c_int32, [_h, _dp, _dp, C.c_int32]),
    "fos_kkt_mul": (C.c_int32, [_h, _dp, _dp]),
    "fos_affine_prox": (C.c_int32, [_h, _dp, _dp]),
    "fos_hsdematrix_prox": (C.c_int32, [_h, _dp, _dp]),
    "fos_cone_prox": (C.c_int32, [_h, _dp, _dp]),
    "fos_cg_dense": (C.c_int32, [_h, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_int64, _i64p]),
    "fos_prox_cone": (C.c_int32, [_h, C.c_int32, C.c_int32, _dp, _dp, C.c_int64]),
    "fos_get_stream": (C.c_int32, [_h, C.POINTER(C.c_uint64)]),
    "fos_k1_plan": (C.c_int32, [C.c_int64, C.c_int64, C.c_int32, _i32p, _i32p, C.c_int64, _i32p, _i32p, C.c_int64]),
    "fos_batch_plan": (C.c_int32, [C.c_int64, C.c_int64, _i64p]),
    "fos_hybrid_plan": (C.c_int32, [C.c_int64, C.c_int64, _i32p, _i64p]),
    "fos_get_tail_trace": (C.c_int32, [_h, _dp]),
    "fos_time_matvec": (C.c_int32, [_h, C.c_int32, C.c_int32, _dp, _dp]),
    "fos_time_psd": (C.c_int32, [_h, C.c_int64, C.c_int64, _dp, _dp, C.c_int32, _dp, _i32p]),
}

_lib = None


def lib_path() -> Path:
    return _build.LIB


def load():
    """Build (if needed) and dlopen the CUDA library; bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    if not Path(path).exists():
        raise RuntimeError(f"{path} is missing and could not be built; fos_b200 has no CPU fallback")
    L = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class FosError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fos_b200 error {code}: {msg}")
        self.code = code


def check(handle, rc):
    if rc != 0:
        msg = load().fos_last_error(handle)
        raise FosError(rc, msg.decode() if msg else "unknown")
