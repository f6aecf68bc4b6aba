"""svdlibrs_b200 -- B200-native LAS2 Lanczos SVD behind the API of the Rust crate dfarnham/svdlibrs.

Host-side mirror of the reference's public interface (reference paths are into its src/lib.rs):

    svd(A)                                   :499   svdLAS2(A, 0, 0, [-1e-30, 1e-30], 1e-6, 0)
    svd_dim(A, dimensions)                   :511
    svd_dim_seed(A, dimensions, random_seed) :524
    svdLAS2(A, dimensions, iterations, end_interval, kappa, random_seed) -> SvdRec   :570-577
    SvdRec {d, ut, s, vt, diagnostics}, SvdRec.recompose()                           :455-461, 2023-2028
    Diagnostics {...11 fields...}                                                    :478-490
    SvdLibError {ImtqlbError, StartvError, StponeError, Imtql2Error, Las2Error}      src/error.rs:8-26

`A` plays the role of nalgebra-sparse's CsrMatrix / CscMatrix / CooMatrix: a scipy.sparse csr / csc / coo
matrix (or anything with the same .format/.shape/.indptr/.indices/.data or .row/.col/.data attributes).
Everything is computed by libsvdlas2_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/svdlas2.h); there is no CPU fallback -- without the library or without a GPU the calls raise.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib

__all__ = ["svd", "svd_dim", "svd_dim_seed", "svdLAS2", "SvdRec", "Diagnostics", "SvdLibError", "Operator", "RawMatrix",
           "ShardedOperator", "Communicator", "comm_unique_id", "build", "trim_caches"]

build = _lib.build

_ERR_KIND = {1: "ImtqlbError", 2: "StartvError", 3: "StponeError", 4: "Imtql2Error", 5: "Las2Error",
             6: "InvalidArgument", 7: "ReferencePanic", 100: "CudaError", 101: "CommError", 102: "AllocError"}
_ERR_PREFIX = {1: "svdlibrs/imtqlb", 2: "svdlibrs/startv", 3: "svdlibrs/stpone", 4: "svdlibrs/imtql2",
               5: "svdlibrs/svdLas2"}


class SvdLibError(Exception):
    """Mirror of the reference's error enum (src/error.rs:8-26).  `.kind` names the variant; codes >= 6 are the
    conditions the Rust crate cannot express (bad raw arrays, a reference `assert!`, CUDA / NCCL failures)."""

    def __init__(self, code: int, message: str):
        self.code = code
        self.kind = _ERR_KIND.get(code, f"Error{code}")
        self.message = message
        prefix = _ERR_PREFIX.get(code, f"svdlas2_b200/{self.kind}")
        super().__init__(f"{prefix}: {message}")

    def __eq__(self, other):  # the reference derives PartialEq
        return isinstance(other, SvdLibError) and (self.code, self.message) == (other.code, other.message)

    __hash__ = Exception.__hash__


@dataclass
class Diagnostics:
    """Computational diagnostics, field for field as src/lib.rs:478-490."""
    non_zero: int
    dimensions: int
    iterations: int
    transposed: bool
    lanczos_steps: int
    ritz_values_stabilized: int
    significant_values: int
    singular_values: int
    end_interval: list
    kappa: float
    random_seed: int


@dataclass
class SvdRec:
    """Singular value decomposition components (src/lib.rs:455-461): ut is d x rows(A), vt is d x cols(A)."""
    d: int
    ut: np.ndarray
    s: np.ndarray
    vt: np.ndarray
    diagnostics: Diagnostics
    profile: dict = field(default_factory=dict, compare=False)
    trace: dict = field(default_factory=dict, compare=False)
    nside_rows: tuple = field(default=(0, 0), compare=False)  # sharded operators: (first, count) of the N-side rows held here

    def recompose(self) -> np.ndarray:
        """ut' * diag(s) * vt  (src/lib.rs:2023-2028)."""
        return self.ut.T @ np.diag(self.s) @ self.vt

    def __eq__(self, other):
        return (isinstance(other, SvdRec) and self.d == other.d and np.array_equal(self.ut, other.ut)
                and np.array_equal(self.s, other.s) and np.array_equal(self.vt, other.vt)
                and self.diagnostics == other.diagnostics)


# ------------------------------------------------------------------------------------------------ helpers

def _u64(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype == np.int64 and a.flags.c_contiguous:
        return a.view(np.uint64)  # non-negative by construction: reinterpret, do not copy
    return np.ascontiguousarray(a, dtype=np.uint64)


class RawMatrix:
    """The raw arrays of a nalgebra-sparse matrix (u64 indices), borrowed without conversion:
    RawMatrix("csr", (nrows, ncols), row_offsets, col_indices, values)   -- CsrMatrix::csr_data()
    RawMatrix("csc", (nrows, ncols), col_offsets, row_indices, values)   -- CscMatrix::csc_data()
    RawMatrix("coo", (nrows, ncols), row_indices, col_indices, values)   -- CooMatrix triplets"""

    def __init__(self, format: str, shape, a, b, data):
        self.format, self.shape, self.data = format, tuple(shape), data
        self.nnz = int(len(data))
        if format == "coo":
            self.row, self.col = a, b
        else:
            self.indptr, self.indices = a, b


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _matrix_arrays(A):
    """(format, nrows, ncols, nnz, a, b, values) with u64 index arrays, as nalgebra-sparse exposes them."""
    fmt = getattr(A, "format", None)
    nrows, ncols = A.shape
    if fmt == "csr":
        return _lib.FMT_CSR, nrows, ncols, int(A.nnz), _u64(A.indptr), _u64(A.indices), _f64(A.data)
    if fmt == "csc":
        return _lib.FMT_CSC, nrows, ncols, int(A.nnz), _u64(A.indptr), _u64(A.indices), _f64(A.data)
    if fmt == "coo":
        return _lib.FMT_COO, nrows, ncols, int(A.nnz), _u64(A.row), _u64(A.col), _f64(A.data)
    raise TypeError("expected a scipy.sparse csr / csc / coo matrix (the CsrMatrix / CscMatrix / CooMatrix of the "
                    f"reference), got {type(A).__name__}")


def _raise(code: int):
    raise SvdLibError(code, _lib.lib().svdlas2_last_error().decode(errors="replace"))


class _ResultOwner:
    """Keeps a svdlas2_result alive while numpy arrays view its (pinned, library-owned) buffers; the buffers go
    back to the library's host pool when the last view is garbage collected (no 0.9 GB memcpy at cfg 2)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            if self.ptr is not None:
                _lib.lib().svdlas2_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _view(owner: _ResultOwner, p, shape) -> np.ndarray:
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape)
    buf = (C.c_double * n).from_address(C.addressof(p.contents))
    buf._owner = owner  # the ctypes array is the numpy base object; it pins the owner
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


def _take_result(ptr) -> SvdRec:
    owner = _ResultOwner(ptr)
    r = ptr.contents
    d = int(r.d)
    # a sharded operator may hand out only a block of the (replicated) N-side vectors per rank: svdlas2_result.nside_row_*
    nb, nc = int(r.nside_row_begin), int(r.nside_row_count)
    n_is_ut = bool(r.diagnostics.transposed)
    ut = _view(owner, r.ut, (nc if n_is_ut else d, int(r.ut_cols)))
    vt = _view(owner, r.vt, (d if n_is_ut else nc, int(r.vt_cols)))
    s = _view(owner, r.s, (d,)).copy()
    g = r.diagnostics
    diag = Diagnostics(int(g.non_zero), int(g.dimensions), int(g.iterations), bool(g.transposed),
                       int(g.lanczos_steps), int(g.ritz_values_stabilized), int(g.significant_values),
                       int(g.singular_values), [g.end_interval[0], g.end_interval[1]], float(g.kappa),
                       int(g.random_seed))
    prof = {name: getattr(r.profile, name) for name, _ in _lib.Profile._fields_}
    n = int(r.trace_len)
    trace = {}
    if n and r.alf:
        trace = dict(alf=_view(owner, r.alf, (n,)).copy(), bet=_view(owner, r.bet, (n + 1,)).copy(),
                     ritz=_view(owner, r.ritz, (n,)).copy(), bnd=_view(owner, r.bnd, (n,)).copy())
    rec = SvdRec(d, ut, s, vt, diag, prof, trace)
    rec.nside_rows = (nb, nc)
    return rec


# ------------------------------------------------------------------------------------------------ public API

def svdLAS2(A, dimensions: int, iterations: int, end_interval: Sequence[float], kappa: float,
            random_seed: int) -> SvdRec:
    """Compute a singular value decomposition (src/lib.rs:570-655); parameters as in the reference."""
    L = _lib.lib()
    fmt, nr, nc, nnz, a, b, v = _matrix_arrays(A)
    fn = {_lib.FMT_CSR: L.svdlas2_csr, _lib.FMT_CSC: L.svdlas2_csc, _lib.FMT_COO: L.svdlas2_coo}[fmt]
    endi = (C.c_double * 2)(float(end_interval[0]), float(end_interval[1]))
    out = C.POINTER(_lib.Result)()
    rc = fn(nr, nc, nnz, _ptr(a, C.c_uint64), _ptr(b, C.c_uint64), _ptr(v, C.c_double), int(dimensions),
            int(iterations), endi, float(kappa), int(random_seed) & 0xFFFFFFFF, C.byref(out))
    if rc != 0:
        _raise(rc)
    return _take_result(out)


def svd(A) -> SvdRec:
    """SVD at full dimensionality (src/lib.rs:499-501)."""
    return svdLAS2(A, 0, 0, (-1.0e-30, 1.0e-30), 1.0e-6, 0)


def svd_dim(A, dimensions: int) -> SvdRec:
    """SVD at desired dimensionality (src/lib.rs:511-513)."""
    return svdLAS2(A, dimensions, 0, (-1.0e-30, 1.0e-30), 1.0e-6, 0)


def svd_dim_seed(A, dimensions: int, random_seed: int) -> SvdRec:
    """SVD at desired dimensionality with supplied seed (src/lib.rs:524-526)."""
    return svdLAS2(A, dimensions, 0, (-1.0e-30, 1.0e-30), 1.0e-6, random_seed)


class Operator:
    """The matrix resident in HBM as CSR(B) + CSR(B') -- the reference's `&dyn SMat` (src/lib.rs:439-444) with the
    upload paid once.  `solve` is svdLAS2; `opa` / `opb` are SMat::svd_opa (:443) and svd_opb (:829-837)."""

    def __init__(self, A, device: int = -1):
        L = _lib.lib()
        fmt, nr, nc, nnz, a, b, v = _matrix_arrays(A)
        self._h = C.c_void_p()
        rc = L.svdlas2_operator_create(fmt, nr, nc, nnz, _ptr(a, C.c_uint64), _ptr(b, C.c_uint64),
                                       _ptr(v, C.c_double), int(device), C.byref(self._h))
        if rc != 0:
            _raise(rc)
        self.shape_a = (nr, nc)
        m, n, z, t = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        L.svdlas2_operator_shape(self._h, C.byref(m), C.byref(n), C.byref(z), C.byref(t))
        self.m, self.n, self.nnz, self.transposed = int(m.value), int(n.value), int(z.value), bool(t.value)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().svdlas2_operator_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set(self, knob: str, value: int):
        rc = _lib.lib().svdlas2_operator_set(self._h, knob.encode(), int(value))
        if rc != 0:
            _raise(rc)

    def info(self, key: str) -> float:
        """Read-only facts about the resident operator (svdlas2_operator_info): "relabelled", "ring", "ring_calls", ..."""
        v = C.c_double()
        rc = _lib.lib().svdlas2_operator_info(self._h, key.encode(), C.byref(v))
        if rc != 0:
            _raise(rc)
        return v.value

    def solve(self, dimensions=0, iterations=0, end_interval=(-1.0e-30, 1.0e-30), kappa=1.0e-6,
              random_seed=0) -> SvdRec:
        endi = (C.c_double * 2)(float(end_interval[0]), float(end_interval[1]))
        out = C.POINTER(_lib.Result)()
        rc = _lib.lib().svdlas2_operator_solve(self._h, int(dimensions), int(iterations), endi, float(kappa),
                                               int(random_seed) & 0xFFFFFFFF, C.byref(out))
        if rc != 0:
            _raise(rc)
        return _take_result(out)

    def opa(self, x, transposed: bool = False) -> np.ndarray:
        nr, nc = self.shape_a
        x = _f64(x)
        if x.shape != ((nr,) if transposed else (nc,)):
            raise ValueError("svd_opa: x has the wrong length")
        y = np.empty(nc if transposed else nr)
        rc = _lib.lib().svdlas2_operator_opa(self._h, _ptr(x, C.c_double), _ptr(y, C.c_double), int(transposed))
        if rc != 0:
            _raise(rc)
        return y

    def opb(self, x) -> np.ndarray:
        x = _f64(x)
        if x.shape != (self.n,):
            raise ValueError("svd_opb: x must have N entries")
        y = np.empty(self.n)
        rc = _lib.lib().svdlas2_operator_opb(self._h, _ptr(x, C.c_double), _ptr(y, C.c_double))
        if rc != 0:
            _raise(rc)
        return y

    def time_opb(self, reps: int = 20, flush_l2: bool = False) -> dict:
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        rc = _lib.lib().svdlas2_operator_time_opb(self._h, int(reps), int(flush_l2), C.byref(a), C.byref(b),
                                                  C.byref(c))
        if rc != 0:
            _raise(rc)
        m, n, z = self.m, self.n, self.nnz
        return dict(ms_per_opb=a.value, ms_spmv_b=b.value, ms_spmv_bt=c.value,
                    bytes_opb=24 * z + 20 * (m + n) + 8,
                    bytes_spmv_b=12 * z + 4 * (m + 1) + 8 * n + 8 * m,
                    bytes_spmv_bt=12 * z + 4 * (n + 1) + 8 * m + 8 * n)


def trim_caches() -> None:
    """Give the cached host result buffers and the cached device scratch back (see svdlas2_trim_caches)."""
    _lib.lib().svdlas2_trim_caches()


def comm_unique_id() -> bytes:
    """Opaque id for ShardedOperator; create on rank 0 and broadcast (torch.distributed, MPI, ...)."""
    buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
    rc = _lib.lib().svdlas2_comm_unique_id(buf)
    if rc != 0:
        _raise(rc)
    return bytes(buf)


class Communicator:
    """Long-lived NCCL communicator of one rank (one process per GPU); serves any number of ShardedOperators."""

    def __init__(self, rank: int, world_size: int, comm_id: Optional[bytes], device: int = -1):
        idbuf = (C.c_uint8 * _lib.COMM_ID_BYTES)(*(comm_id or bytes(_lib.COMM_ID_BYTES)))
        self._h = C.c_void_p()
        rc = _lib.lib().svdlas2_comm_create(int(rank), int(world_size), idbuf, int(device), C.byref(self._h))
        if rc != 0:
            _raise(rc)
        self.rank, self.world_size = rank, world_size

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:  # (_lib is None once the interpreter tears modules down)
            _lib.lib().svdlas2_comm_destroy(self._h)
            self._h = None

    __del__ = close


class ShardedOperator(Operator):
    """One rank's row range [row_begin, row_end) of the oriented operator B: rows are partitioned by
    nnz over the GPUs of one box, y = sum_g B_g'(B_g x) is completed by one all-reduce of N doubles per opb."""

    def __init__(self, local_csr, m_global: int, row_begin: int, global_nnz: int, comm: Optional["Communicator"],
                 b_is_a_transposed: bool = False, device: int = -1):
        L = _lib.lib()
        if getattr(local_csr, "format", None) != "csr":
            raise TypeError("the shard must be a csr matrix holding rows [row_begin, row_end) of B")
        rows, n = local_csr.shape
        a, b, v = _u64(local_csr.indptr), _u64(local_csr.indices), _f64(local_csr.data)
        self._h = C.c_void_p()
        self._comm = comm  # keep the communicator alive as long as the operator
        rc = L.svdlas2_operator_create_shard(int(m_global), int(n), int(row_begin), int(row_begin + rows),
                                             int(local_csr.nnz), _ptr(a, C.c_uint64), _ptr(b, C.c_uint64),
                                             _ptr(v, C.c_double), int(b_is_a_transposed), int(global_nnz),
                                             comm._h if comm is not None else None, int(device), C.byref(self._h))
        if rc != 0:
            _raise(rc)
        self.shape_a = (n, m_global) if b_is_a_transposed else (m_global, n)
        self.m, self.n, self.nnz, self.transposed = rows, n, int(local_csr.nnz), bool(b_is_a_transposed)
        self.m_global, self.row_begin = m_global, row_begin
        self.rank = comm.rank if comm else 0
        self.world_size = comm.world_size if comm else 1

    def opa(self, x, transposed: bool = False):
        raise SvdLibError(6, "opa is not available on a sharded operator")
