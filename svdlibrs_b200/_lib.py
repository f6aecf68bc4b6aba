"""ctypes binding of libsvdlas2_b200.so (include/svdlas2.h).  No fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "libsvdlas2_b200.so")

COMM_ID_BYTES = 128
FMT_CSR, FMT_CSC, FMT_COO = 0, 1, 2

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class Diagnostics(C.Structure):
    _fields_ = [
        ("non_zero", C.c_uint64), ("dimensions", C.c_uint64), ("iterations", C.c_uint64),
        ("transposed", C.c_int32),
        ("lanczos_steps", C.c_uint64), ("ritz_values_stabilized", C.c_uint64),
        ("significant_values", C.c_uint64), ("singular_values", C.c_uint64),
        ("end_interval", C.c_double * 2), ("kappa", C.c_double), ("random_seed", C.c_uint32),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("t_total", C.c_double), ("t_upload", C.c_double), ("t_lanczos", C.c_double), ("t_ritvec", C.c_double),
        ("t_imtql2", C.c_double), ("t_download", C.c_double), ("t_opb_device", C.c_double), ("t_device", C.c_double), ("t_sync_wait", C.c_double),
        ("n_opb", C.c_uint64), ("n_spmv", C.c_uint64), ("n_purge_fired", C.c_uint64),
        ("n_purge_passes", C.c_uint64), ("n_purge_vectors", C.c_uint64), ("n_restarts", C.c_uint64),
        ("n_kernel_launches", C.c_uint64), ("n_host_syncs", C.c_uint64),
        ("opb_algorithmic_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
    ]


class Result(C.Structure):
    _fields_ = [
        ("d", C.c_uint64), ("ut_cols", C.c_uint64), ("vt_cols", C.c_uint64),
        ("ut", f64p), ("s", f64p), ("vt", f64p),
        ("diagnostics", Diagnostics), ("profile", Profile),
        ("trace_len", C.c_uint64), ("alf", f64p), ("bet", f64p), ("ritz", f64p), ("bnd", f64p),
        ("nside_row_begin", C.c_uint64), ("nside_row_count", C.c_uint64),
    ]


# every symbol include/svdlas2.h declares (tests/test_abi_symbols.py checks the header against this list)
SYMBOLS = [
    "svdlas2_csr", "svdlas2_csc", "svdlas2_coo", "svdlas2_free", "svdlas2_last_error", "svdlas2_abi_version",
    "svdlas2_trim_caches",
    "svdlas2_operator_create", "svdlas2_operator_destroy", "svdlas2_operator_solve", "svdlas2_operator_opa",
    "svdlas2_operator_opb", "svdlas2_operator_shape", "svdlas2_operator_time_opb", "svdlas2_operator_set",
    "svdlas2_operator_info",
    "svdlas2_comm_unique_id", "svdlas2_comm_create", "svdlas2_comm_destroy", "svdlas2_operator_create_shard",
]


def build(force: bool = False, jobs: int = 8) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", _CSRC, f"-j{jobs}"]
    if force:
        cmd.append("-B")
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library.  Raises ImportError (never falls back to a CPU path) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C svdlibrs_b200/csrc` (or __graft_entry__.build()). "
            "svdlibrs_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    RP = C.POINTER(C.POINTER(Result))
    one_shot = [C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, C.c_uint64, C.c_uint64, f64p, C.c_double,
                C.c_uint32, RP]
    for name in ("svdlas2_csr", "svdlas2_csc", "svdlas2_coo"):
        getattr(L, name).argtypes = one_shot
        getattr(L, name).restype = C.c_int
    L.svdlas2_free.argtypes = [C.POINTER(Result)]
    L.svdlas2_free.restype = None
    L.svdlas2_last_error.argtypes = []
    L.svdlas2_last_error.restype = C.c_char_p
    L.svdlas2_abi_version.argtypes = []
    L.svdlas2_trim_caches.argtypes = []
    L.svdlas2_trim_caches.restype = C.c_int
    L.svdlas2_abi_version.restype = C.c_int
    OP = C.c_void_p
    L.svdlas2_operator_create.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, C.c_int,
                                          C.POINTER(OP)]
    L.svdlas2_operator_create.restype = C.c_int
    L.svdlas2_operator_destroy.argtypes = [OP]
    L.svdlas2_operator_destroy.restype = None
    L.svdlas2_operator_solve.argtypes = [OP, C.c_uint64, C.c_uint64, f64p, C.c_double, C.c_uint32, RP]
    L.svdlas2_operator_solve.restype = C.c_int
    L.svdlas2_operator_opa.argtypes = [OP, f64p, f64p, C.c_int]
    L.svdlas2_operator_opa.restype = C.c_int
    L.svdlas2_operator_opb.argtypes = [OP, f64p, f64p]
    L.svdlas2_operator_opb.restype = C.c_int
    L.svdlas2_operator_shape.argtypes = [OP, u64p, u64p, u64p, C.POINTER(C.c_int)]
    L.svdlas2_operator_shape.restype = C.c_int
    L.svdlas2_operator_time_opb.argtypes = [OP, C.c_int, C.c_int, f64p, f64p, f64p]
    L.svdlas2_operator_time_opb.restype = C.c_int
    L.svdlas2_operator_set.argtypes = [OP, C.c_char_p, C.c_int64]
    L.svdlas2_operator_set.restype = C.c_int
    L.svdlas2_operator_info.argtypes = [OP, C.c_char_p, C.POINTER(C.c_double)]
    L.svdlas2_operator_info.restype = C.c_int
    L.svdlas2_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.svdlas2_comm_unique_id.restype = C.c_int
    L.svdlas2_comm_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_void_p)]
    L.svdlas2_comm_create.restype = C.c_int
    L.svdlas2_comm_destroy.argtypes = [C.c_void_p]
    L.svdlas2_comm_destroy.restype = None
    L.svdlas2_operator_create_shard.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p,
                                                f64p, C.c_int, C.c_uint64, C.c_void_p, C.c_int, C.POINTER(OP)]
    L.svdlas2_operator_create_shard.restype = C.c_int
    _lib = L
    return L
