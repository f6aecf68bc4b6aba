"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of the CPU oracle (oracle/las2_oracle.c, a restatement of the reference's
src/lib.rs:570-2130).  May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblas2_oracle.so")

FMT_CSR, FMT_CSC, FMT_COO = 0, 1, 2
ERR_NAMES = {0: "ok", 1: "ImtqlbError", 2: "StartvError", 3: "StponeError", 4: "Imtql2Error", 5: "Las2Error",
             7: "Panic"}


class _Result(C.Structure):
    _fields_ = [
        ("d", C.c_uint64), ("ut_cols", C.c_uint64), ("vt_cols", C.c_uint64),
        ("ut", C.POINTER(C.c_double)), ("s", C.POINTER(C.c_double)), ("vt", C.POINTER(C.c_double)),
        ("non_zero", C.c_uint64), ("dimensions", C.c_uint64), ("iterations", C.c_uint64),
        ("transposed", C.c_int),
        ("lanczos_steps", C.c_uint64), ("ritz_values_stabilized", C.c_uint64),
        ("significant_values", C.c_uint64), ("singular_values", C.c_uint64),
        ("end_interval", C.c_double * 2), ("kappa", C.c_double), ("random_seed", C.c_uint32),
        ("trace_len", C.c_uint64),
        ("alf", C.POINTER(C.c_double)), ("bet", C.POINTER(C.c_double)),
        ("ritz", C.POINTER(C.c_double)), ("bnd", C.POINTER(C.c_double)),
        ("n_opb", C.c_uint64), ("n_opa", C.c_uint64),
        ("n_purge_fired", C.c_uint64), ("n_purge_vecs", C.c_uint64), ("n_restarts", C.c_uint64),
        ("t_total", C.c_double), ("t_opb", C.c_double), ("t_purge", C.c_double),
        ("t_imtql2", C.c_double), ("t_ritvec", C.c_double),
        ("message", C.c_char * 160),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/liblas2_oracle.so with gcc (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("las2_oracle.c", "chacha12.c", "las2_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
        L.las2_oracle_svd.restype = C.c_int
        L.las2_oracle_svd.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, C.c_uint64,
                                      C.c_uint64, f64p, C.c_double, C.c_uint32, C.POINTER(C.POINTER(_Result))]
        L.las2_oracle_free.argtypes = [C.POINTER(_Result)]
        L.las2_oracle_free.restype = None
        L.las2_oracle_opa.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, f64p, f64p,
                                      C.c_int]
        L.las2_oracle_opa.restype = None
        L.las2_oracle_opb.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u64p, f64p, f64p, f64p,
                                      f64p, C.c_int]
        L.las2_oracle_opb.restype = None
        L.las2_oracle_unit_costs.argtypes = [C.c_uint64, C.c_uint64, f64p]
        L.las2_oracle_unit_costs.restype = None
        L.las2_oracle_imtqlb.argtypes = [C.c_uint64, f64p, f64p, f64p]
        L.las2_oracle_imtqlb.restype = C.c_int
        L.las2_oracle_imtql2.argtypes = [C.c_uint64, C.c_uint64, f64p, f64p, f64p]
        L.las2_oracle_imtql2.restype = C.c_int
        L.las2_oracle_pythag.argtypes = [C.c_double, C.c_double]
        L.las2_oracle_pythag.restype = C.c_double
        L.las2_oracle_rotate.argtypes = [f64p, C.c_uint64, C.c_uint64]
        L.las2_oracle_rotate.restype = None
        L.oracle_start_vector.argtypes = [C.c_uint32, C.c_size_t, f64p]
        L.oracle_start_vector.restype = None
        L.oracle_chacha_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_int, C.POINTER(C.c_uint32)]
        L.oracle_chacha_block.restype = None
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.kind = ERR_NAMES.get(code, str(code))
        self.message = message


@dataclass
class OracleSvd:
    d: int
    ut: np.ndarray
    s: np.ndarray
    vt: np.ndarray
    diagnostics: dict
    trace: dict = field(default_factory=dict)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def matrix_arrays(A):
    """scipy.sparse matrix -> (fmt, nrows, ncols, nnz, a, b, values) with u64 indices, like nalgebra-sparse."""
    fmt = A.format
    if fmt == "csr":
        return FMT_CSR, A.shape[0], A.shape[1], A.nnz, _u64(A.indptr), _u64(A.indices), _f64(A.data)
    if fmt == "csc":
        return FMT_CSC, A.shape[0], A.shape[1], A.nnz, _u64(A.indptr), _u64(A.indices), _f64(A.data)
    if fmt == "coo":
        return FMT_COO, A.shape[0], A.shape[1], A.nnz, _u64(A.row), _u64(A.col), _f64(A.data)
    raise TypeError(f"unsupported sparse format {fmt!r}")


def svd_las2(A, dimensions=0, iterations=0, end_interval=(-1.0e-30, 1.0e-30), kappa=1.0e-6, random_seed=0):
    """Oracle svdLAS2 (reference src/lib.rs:570-655) on a scipy.sparse csr/csc/coo matrix."""
    L = lib()
    fmt, nr, nc, nnz, a, b, v = matrix_arrays(A)
    endi = (C.c_double * 2)(*end_interval)
    out = C.POINTER(_Result)()
    rc = L.las2_oracle_svd(fmt, nr, nc, nnz, _p(a, C.c_uint64), _p(b, C.c_uint64), _p(v, C.c_double),
                           int(dimensions), int(iterations), endi, float(kappa), int(random_seed), C.byref(out))
    try:
        r = out.contents
        if rc != 0:
            raise OracleError(rc, r.message.decode())
        d = int(r.d)
        ut = np.ctypeslib.as_array(r.ut, shape=(d, int(r.ut_cols))).copy() if d else np.zeros((0, int(r.ut_cols)))
        vt = np.ctypeslib.as_array(r.vt, shape=(d, int(r.vt_cols))).copy() if d else np.zeros((0, int(r.vt_cols)))
        s = np.ctypeslib.as_array(r.s, shape=(d,)).copy() if d else np.zeros(0)
        n = int(r.trace_len)
        diag = dict(non_zero=int(r.non_zero), dimensions=int(r.dimensions), iterations=int(r.iterations),
                    transposed=bool(r.transposed), lanczos_steps=int(r.lanczos_steps),
                    ritz_values_stabilized=int(r.ritz_values_stabilized),
                    significant_values=int(r.significant_values), singular_values=int(r.singular_values),
                    end_interval=[r.end_interval[0], r.end_interval[1]], kappa=float(r.kappa),
                    random_seed=int(r.random_seed))
        trace = dict(alf=np.ctypeslib.as_array(r.alf, shape=(n,)).copy(),
                     bet=np.ctypeslib.as_array(r.bet, shape=(n + 1,)).copy(),
                     ritz=np.ctypeslib.as_array(r.ritz, shape=(n,)).copy(),
                     bnd=np.ctypeslib.as_array(r.bnd, shape=(n,)).copy(),
                     n_opb=int(r.n_opb), n_opa=int(r.n_opa), n_purge_fired=int(r.n_purge_fired),
                     n_purge_vecs=int(r.n_purge_vecs), n_restarts=int(r.n_restarts),
                     t_total=r.t_total, t_opb=r.t_opb, t_purge=r.t_purge, t_imtql2=r.t_imtql2,
                     t_ritvec=r.t_ritvec)
        return OracleSvd(d, ut, s, vt, diag, trace)
    finally:
        L.las2_oracle_free(out)


def svd(A):
    return svd_las2(A, 0, 0, (-1.0e-30, 1.0e-30), 1.0e-6, 0)


def svd_dim(A, dimensions):
    return svd_las2(A, dimensions, 0, (-1.0e-30, 1.0e-30), 1.0e-6, 0)


def svd_dim_seed(A, dimensions, random_seed):
    return svd_las2(A, dimensions, 0, (-1.0e-30, 1.0e-30), 1.0e-6, random_seed)


def opa(A, x, transposed=False):
    L = lib()
    fmt, nr, nc, nnz, a, b, v = matrix_arrays(A)
    x = _f64(x)
    y = np.zeros(nc if transposed else nr)
    L.las2_oracle_opa(fmt, nr, nc, nnz, _p(a, C.c_uint64), _p(b, C.c_uint64), _p(v, C.c_double),
                      _p(x, C.c_double), _p(y, C.c_double), int(transposed))
    return y


def opb(A, x, transposed=False):
    """y = B'(B x) with B = A (transposed False) or A' (True); reference src/lib.rs:829-837."""
    L = lib()
    fmt, nr, nc, nnz, a, b, v = matrix_arrays(A)
    x = _f64(x)
    n = nr if transposed else nc
    m = nc if transposed else nr
    y = np.zeros(n)
    tmp = np.zeros(m)
    L.las2_oracle_opb(fmt, nr, nc, nnz, _p(a, C.c_uint64), _p(b, C.c_uint64), _p(v, C.c_double),
                      _p(x, C.c_double), _p(y, C.c_double), _p(tmp, C.c_double), int(transposed))
    return y


def unit_costs(n: int, nvec: int = 8) -> dict:
    """Seconds per purge visit of a stored vector, per lanczos_step BLAS-1 chain, per Ritz-contraction term (vectors of
    length n) on this host: the O(N) pieces of bench.py's bounded CPU sample."""
    out = (C.c_double * 3)()
    lib().las2_oracle_unit_costs(int(n), int(nvec), out)
    return dict(purge_vec_s=out[0], step_blas1_s=out[1], contract_term_s=out[2])


def opb_raw(fmt: int, nrows: int, ncols: int, a: np.ndarray, b: np.ndarray, v: np.ndarray, x: np.ndarray, transposed: bool):
    """svd_opb on raw u64 arrays without any conversion copy (the 15 GB of cfg 4)."""
    L = lib()
    n, m = (nrows, ncols) if transposed else (ncols, nrows)
    y, tmp = np.zeros(n), np.zeros(m)
    L.las2_oracle_opb(fmt, nrows, ncols, len(v), _p(a, C.c_uint64), _p(b, C.c_uint64), _p(v, C.c_double),
                      _p(x, C.c_double), _p(y, C.c_double), _p(tmp, C.c_double), int(transposed))
    return y


def start_vector(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    lib().oracle_start_vector(int(seed), n, _p(out, C.c_double))
    return out


def chacha_block(key_words, counter: int, rounds: int) -> np.ndarray:
    key = (C.c_uint32 * 8)(*[int(k) for k in key_words])
    out = (C.c_uint32 * 16)()
    lib().oracle_chacha_block(key, int(counter), int(rounds), out)
    return np.array(list(out), dtype=np.uint32)


def imtqlb(d, e, bnd):
    d, e, bnd = _f64(d).copy(), _f64(e).copy(), _f64(bnd).copy()
    rc = lib().las2_oracle_imtqlb(len(d), _p(d, C.c_double), _p(e, C.c_double), _p(bnd, C.c_double))
    return rc, d, e, bnd


def imtql2(d, e):
    d, e = _f64(d).copy(), _f64(e).copy()
    n = len(d)
    z = np.eye(n)
    rc = lib().las2_oracle_imtql2(n, n, _p(d, C.c_double), _p(e, C.c_double), _p(z, C.c_double))
    return rc, d, z


def pythag(a, b):
    return lib().las2_oracle_pythag(float(a), float(b))


def rotate(a, x):
    a = _f64(a).copy()
    lib().las2_oracle_rotate(_p(a, C.c_double), len(a), int(x))
    return a
