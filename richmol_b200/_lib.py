"""ctypes binding of librichmol_b200.so (the C ABI in include/richmol_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C richmol_b200/csrc`.
Loading failures are loud: the product path has no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RMB_LIB") or os.path.join(_HERE, "librichmol_b200.so")   # RMB_LIB: A/B builds (tools/)

RMB_OK, RMB_ERR_INVALID, RMB_ERR_CUDA, RMB_ERR_MAXORDER, RMB_ERR_NOFIELD = 0, -1, -2, -3, -4

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)


class PartDesc(C.Structure):
    _fields_ = [
        ("ncart", C.c_int32), ("nprod", C.c_int32),
        ("pr_bra", c_i32p), ("pr_ket", c_i32p), ("pr_table", c_i32p), ("pr_koff", c_i64p),
        ("k_is_complex", C.c_int32), ("kpool", c_f64p), ("kpool_len", C.c_int64),
        ("ntables", C.c_int32), ("tb_dm1", c_i32p), ("tb_dm2", c_i32p), ("tb_nd", c_i32p),
        ("tb_off", c_i64p), ("ent_col", c_i32p), ("ent_coef", c_f64p),
    ]


class OperatorDesc(C.Structure):
    _fields_ = [
        ("nblocks", C.c_int32), ("blk_off", c_i64p), ("blk_dm", c_i32p), ("blk_dk", c_i32p),
        ("nparts", C.c_int32), ("parts", C.POINTER(PartDesc)),
    ]


# name -> (restype, argtypes); every symbol declared in include/richmol_b200.h
SYMBOLS = {
    "rmb_abi_version": (C.c_int32, []),
    "rmb_last_error": (C.c_char_p, []),
    "rmb_device_count": (C.c_int32, []),
    "rmb_operator_create": (C.c_int32, [C.POINTER(OperatorDesc), C.POINTER(C.c_void_p)]),
    "rmb_operator_destroy": (None, [C.c_void_p]),
    "rmb_operator_dim": (C.c_int64, [C.c_void_p]),
    "rmb_operator_nentries": (C.c_int64, [C.c_void_p, C.c_int32]),
    "rmb_operator_set_field": (C.c_int32, [C.c_void_p, C.c_int32, c_f64p, C.c_double, C.c_int32, C.c_void_p]),
    "rmb_operator_get_mf": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "rmb_operator_mf_nonempty": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "rmb_matvec": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "rmb_propagate_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                       C.c_double, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "rmb_propagate_step_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                            C.c_double, C.c_double, C.c_double, C.c_int32, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_void_p]),
    "rmb_propagate_step_host_obs": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                                C.c_double, C.c_double, C.c_double, C.c_int32, C.c_void_p,
                                                C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p),
                                                C.c_void_p, C.c_void_p]),
    "rmb_propagate_many": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_double,
                                       C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_int32, c_i32p, c_f64p,
                                       c_f64p, c_i32p, C.c_int32, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "rmb_expectation": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "rmb_populations": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "rmb_set_workspace_budget": (C.c_int32, [C.c_void_p, C.c_int64]),
    "rmb_get_counters": (C.c_int32, [C.c_void_p, c_i64p]),
    "rmb_operator_work": (C.c_int32, [C.c_void_p, c_f64p, c_f64p, C.c_void_p]),
    "rmb_matvec_timing": (C.c_int32, [C.c_void_p, C.c_int32, c_f64p, c_i64p]),
    "rmb_operator_info": (C.c_int32, [C.c_void_p, c_i64p]),
    "rmb_fp64_peak": (C.c_int32, [c_f64p, c_f64p, C.c_void_p]),
    "rmb_small_expm": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                                   C.c_void_p]),
    "rmb_threej_band": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p,
                                    C.c_void_p]),
}

_lib = None


class LibraryError(RuntimeError):
    pass


def lib():
    """Returns the loaded shared library (raises LibraryError if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C richmol_b200/csrc` (there is no CPU fallback for the TDSE hot path)")
        try:
            l = C.CDLL(LIB_PATH)
        except OSError as e:
            raise LibraryError(f"cannot load {LIB_PATH}: {e}") from None
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        if l.rmb_abi_version() != 1:
            raise LibraryError("librichmol_b200.so ABI version mismatch")
        _lib = l
    return _lib


def last_error():
    return lib().rmb_last_error().decode("utf-8", "replace")


def check(status):
    """Maps C status codes to the exceptions of the reference API."""
    if status == RMB_OK:
        return
    msg = last_error()
    if status == RMB_ERR_MAXORDER:
        raise ValueError(msg)
    if status == RMB_ERR_NOFIELD:
        raise AttributeError(msg)
    if status == RMB_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def require_device():
    n = lib().rmb_device_count()
    if n <= 0:
        raise RuntimeError(
            "richmol_b200 needs a CUDA device (B200, sm_100a); none is visible"
            + (f": {last_error()}" if n < 0 else "") + " -- there is no CPU fallback")
    return n


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
