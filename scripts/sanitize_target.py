"""compute-sanitizer target: every hand-written kernel family once, at sizes that finish in seconds under the tool.
   compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_target.py
Covers: upload (radix transposition, relabelling), k_spmv_chunk in its plain / hot-prefix / slab-local / wide-slab modes and the
heavy / light split, k_spmv_stream (small-matrix path), k_spmm4_chunk, BLAS-1 chain, panel reorthogonalisation, QL replay,
both Ritz contraction kernels, the peer-ring kernel (emulated ranks)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import svdlibrs_b200 as las2
from svdlibrs_b200 import _lib, synthetic

L = _lib.lib()
f64p = C.POINTER(C.c_double)


def solve(A, dims, env):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        r = las2.svd_dim_seed(A, dims, 12345)
        print(f"ok {env} shape={A.shape} nnz={A.nnz} d={r.d} steps={r.diagnostics.lanczos_steps} sigma0={r.s[0]:.6g}", flush=True)
    finally:
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v


A, _ = synthetic.make("cfg2", scale=0.004)      # 4000 x 400, ~1.3e5 nnz
A3, _ = synthetic.make("cfg3", scale=0.002)     # wide CSC, transposed path
solve(A, 8, {})                                                                  # small-matrix kernels
solve(A, 8, {"SVDLAS2_SPMV": "chunk", "SVDLAS2_RELABEL": "1"})                   # chunk stream, hot prefix, relabelled
solve(A3, 6, {"SVDLAS2_SPMV": "chunk", "SVDLAS2_SLABS": "1"})                    # slab-local mode
solve(A3, 6, {"SVDLAS2_SPMV": "chunk", "SVDLAS2_WIDE_SLABS": "1", "SVDLAS2_RELABEL": "1"})
solve(A, 8, {"SVDLAS2_SPMV": "chunk", "SVDLAS2_RITZ": "fma", "SVDLAS2_NO_SPMM": "1"})
# ring kernel, emulated ranks, both forms
L.svdlas2_test_ring_emulation.argtypes = [C.c_int, C.c_uint64, f64p, f64p, C.c_double, f64p, C.c_int, C.c_int, f64p, f64p]
rng = np.random.default_rng(1)
for two_shot in (0, 1):
    G, n = 4, 32 * 257
    ys, qp, q = rng.standard_normal((G, n)), rng.standard_normal(n), rng.standard_normal(n)
    out, dots = np.empty((G, n)), np.empty(G)
    p = lambda a: a.ctypes.data_as(f64p)
    rc = L.svdlas2_test_ring_emulation(G, n, p(ys), p(qp), 0.5, p(q), two_shot, 8, p(out), p(dots))
    print("ring emulation", two_shot, rc, float(np.abs(out[0] - out[3]).max()), flush=True)
# Ritz kernels on a ragged shape
L.svdlas2_test_time_ritz.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, f64p, f64p, f64p]
ms, d1, d2 = (C.c_double * 2)(), C.c_double(), C.c_double()
print("ritz", L.svdlas2_test_time_ritz(3001, 75, 21, 1, -1, ms, C.byref(d1), C.byref(d2)), d1.value, d2.value, flush=True)
print("SANITIZE_TARGET_DONE", flush=True)
