"""
ctypes binding of libdrcvar.so (C ABI declared in include/drcvar.h).

The reference has no FFI layer for this path (it calls cvxpy -> ECOS from Python,
core/risk_metrics.py:156,244); this stub is what replaces that call.  The product path has no
CPU fallback: if the shared library is missing, importing the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DRCVAR_LIB", os.path.join(_HERE, "libdrcvar.so"))

HOST = -1
OK = 0
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM = -1, -2, -3, -4
FLAG_SYNC, FLAG_GENERAL_ONLY, FLAG_NO_BULK, FLAG_FORCE_STREAMING, FLAG_NO_CLUSTER, FLAG_FORCE_CLUSTER = 1, 2, 4, 8, 16, 32
FLAG_NO_PIPELINE = 64
FLAG_LARGE_COORDS = 128   # fp32 samples far from the origin: first-sample-relative sums in the pipelined kernel (speed only)
STATUS_NONFINITE, STATUS_GENERAL, STATUS_DEGENERATE = 1, 2, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

_HALFSPACE_TAIL = [
    C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,      # B, N, stride_b, stride_n, stride_c
    C.c_void_p, C.c_void_p,                                       # ego, h_in
    C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,   # alpha, delta, epsilon, r_robot, r_obs
    C.c_uint32,                                                   # flags
    C.c_void_p, C.c_void_p, C.c_void_p,                           # h_out, h_mean_out, g_out
    C.c_void_p, C.c_void_p, C.c_void_p,                           # cvar_out, var_out, gstar_out
    C.c_void_p, C.c_void_p,                                       # status_out, tail_idx_out
    C.c_int, C.c_void_p,                                          # device, stream
]

# every symbol include/drcvar.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "drcvar_version": (C.c_int, []),
    "drcvar_last_error": (C.c_char_p, []),
    "drcvar_device_count": (C.c_int, []),
    "drcvar_reduction_lanes": (C.c_int, []),
    "drcvar_tail_count": (C.c_int64, [C.c_double, C.c_int64, _dp]),
    "drcvar_max_samples": (C.c_int64, [C.c_int, C.c_int]),
    "drcvar_halfspaces_f32": (C.c_int, [C.c_void_p] + _HALFSPACE_TAIL),
    "drcvar_halfspaces_f64": (C.c_int, [C.c_void_p] + _HALFSPACE_TAIL),
    "drcvar_halfspaces_generated_f32": (C.c_int, [
        C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64,     # mean, chol, seed, index_offset, B, N
        C.c_void_p, C.c_void_p,                                                   # ego, h_in
        C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32,   # alpha .. r_obs, flags
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,   # h, h_mean, g, cvar, var, gstar
        C.c_void_p, C.c_void_p, C.c_void_p,                                       # status, tail_idx, samples_out
        C.c_int, C.c_void_p]),                                                    # device, stream
    "drcvar_trajectory_f64": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drcvar_host_alloc": (C.c_void_p, [C.c_size_t]),
    "drcvar_host_free": (None, [C.c_void_p]),
    "drcvar_launch_count": (C.c_int64, []),
    "drcvar_debug_check_failures": (C.c_int64, [_ip]),
    "drcvar_cluster_ctas": (C.c_int, [C.c_int64, C.c_int, C.c_int64]),
    "drcvar_last_host_call_stats": (C.c_int, [_dp, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}

_lib = None


class DrcvarError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdrcvar error {code}: {msg}")
        self.code = code


def load():
    """Load libdrcvar.so (once) and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()' "
            f"or make -C {os.path.join(_HERE, 'csrc')}).  There is no CPU fallback on the product path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        raise DrcvarError(rc, load().drcvar_last_error().decode("utf-8", "replace"))
    return rc
