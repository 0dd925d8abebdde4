"""ctypes binding of libqbx.so (include/qbx.h).  There is no CPU fallback: a missing
library or a missing GPU raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqbx.so")

_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)

# name -> (argtypes); every function returns int (include/qbx.h)
SIGNATURES = {
    "qbx_init": [C.c_int, C.POINTER(C.c_int)],
    "qbx_shutdown": [],
    "qbx_basis_create": [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                         C.c_void_p, C.POINTER(C.c_void_p)],
    "qbx_basis_destroy": [C.c_void_p],
    "qbx_basis_info": [C.c_void_p, C.c_void_p],
    "qbx_eri_tensor": [C.c_void_p, C.c_void_p, C.c_int64],
    "qbx_eri_quartets": [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
    "qbx_eri_store": [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int],
    "qbx_eri_recompute": [C.c_void_p],
    "qbx_eri_recompute_async": [C.c_void_p],
    "qbx_set_stream": [C.c_void_p],
    "qbx_class_stats": [C.c_void_p, C.c_void_p],
    "qbx_fp64_peak": [C.c_void_p],
    "qbx_pool_trim": [C.c_void_p],
    "qbx_fock_build": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    "qbx_fock_build_device": [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "qbx_one_body": [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p],
    "qbx_boys": [C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "qbx_prim_batch": [C.c_int] * 5 + [C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p],
    "qbx_stats": [C.c_void_p, C.c_void_p, C.c_int],
    "qbx_comm_unique_id": [C.c_void_p],
    "qbx_comm_init": [C.c_int, C.c_int, C.c_void_p],
    "qbx_comm_info": [C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "qbx_comm_destroy": [],
    "qbx_scf_create": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)],
    "qbx_scf_destroy": [C.c_void_p],
    "qbx_scf_set": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "qbx_scf_get": [C.c_void_p, C.c_int, C.c_int, C.c_void_p],
    "qbx_scf_step": [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p],
    "qbx_scf_hist_store": [C.c_void_p, C.c_int],
    "qbx_scf_hist_gram": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    "qbx_scf_combine": [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    "qbx_mo_transform": [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64],
    "qbx_mo_coulomb_ab": [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
}


class QbxError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen libqbx.so and declare every symbol of include/qbx.h (no device is touched)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QbxError(f"{LIB_PATH} is missing: build it with `python quiqbox.jl_b200/build.py` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        lib.qbx_last_error.restype = C.c_char_p
        lib.qbx_last_error.argtypes = []
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise QbxError(f"qbx error {rc}: {load().qbx_last_error().decode()}")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def init(device=None):
    """Bind this process to a GPU (LOCAL_RANK under torchrun, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    n = C.c_int(0)
    check(load().qbx_init(int(device), C.byref(n)))
    return n.value
