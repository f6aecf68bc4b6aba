"""Run the CPU oracle ONCE, at full size, on BASELINE configs that are too big to re-run inside a test, and record what the
GPU tests need to assert north-star parity: every Diagnostics count, all singular values, the head of the alf/bet trace,
and a sketch of U and V (sampled coordinates + projections on fixed pseudo-random directions).

    python scripts/make_oracle_fixtures.py cfg4 cfg5        # hours, single thread like the reference; run in the background
    -> tests/golden/oracle_<cfg>.json   (one file per config, written when that config finishes)

The matrix comes from svdlibrs_b200.synthetic.make(name) (matrix seed 12345), the solve is svd_dim_seed(A, dims, 12345).
TEST INFRASTRUCTURE: this script and the oracle never run on a product path.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from svdlibrs_b200 import synthetic  # noqa: E402

N_SAMPLES, N_PROJ, SKETCH_SEED = 32, 4, 777


def sketch_plan(length: int, tag: int):
    """Sampled coordinates and projection directions for vectors of `length` (shared with tests/helpers.py)."""
    rng = np.random.default_rng([SKETCH_SEED, tag, length])
    idx = np.sort(rng.choice(length, size=min(N_SAMPLES, length), replace=False))
    omega = rng.standard_normal((N_PROJ, length))
    return idx, omega


def run(name: str):
    t0 = time.time()
    A, dims = synthetic.make(name)
    gen_s = time.time() - t0
    print(f"[{name}] generated {A.shape} nnz={A.nnz} in {gen_s:.0f} s", flush=True)
    L = oracle.lib()
    fmt = {"csr": oracle.FMT_CSR, "csc": oracle.FMT_CSC}[A.format]
    # int64 -> uint64 views: no second copy of the 7 GB index array
    a = np.ascontiguousarray(A.indptr).astype(np.uint64)
    b = np.ascontiguousarray(A.indices)
    b = b.view(np.uint64) if b.dtype == np.int64 else b.astype(np.uint64)
    v = np.ascontiguousarray(A.data, dtype=np.float64)
    nr, nc, nnz = A.shape[0], A.shape[1], int(A.nnz)
    lane = np.diff(A.indptr)
    shape_info = dict(shape=[nr, nc], nnz=nnz, format=A.format, lane_len_max=int(lane.max()), lane_len_mean=float(lane.mean()),
                      values_sum=float(A.data.sum()), indices_sum=int(A.indices.sum(dtype=np.int64)))
    del A
    endi = (C.c_double * 2)(-1.0e-30, 1.0e-30)
    out = C.POINTER(oracle._Result)()
    t0 = time.time()
    rc = L.las2_oracle_svd(fmt, nr, nc, nnz, oracle._p(a, C.c_uint64), oracle._p(b, C.c_uint64), oracle._p(v, C.c_double),
                           int(dims), 0, endi, 1.0e-6, 12345, C.byref(out))
    wall = time.time() - t0
    r = out.contents
    if rc != 0:
        raise SystemExit(f"[{name}] oracle failed: {rc} {r.message.decode()}")
    d, n = int(r.d), int(r.trace_len)
    ut = np.ctypeslib.as_array(r.ut, shape=(d, int(r.ut_cols)))  # views, no copies (16 GB at cfg 4)
    vt = np.ctypeslib.as_array(r.vt, shape=(d, int(r.vt_cols)))
    s = np.ctypeslib.as_array(r.s, shape=(d,)).copy()
    iu, ou = sketch_plan(int(r.ut_cols), 1)
    iv, ov = sketch_plan(int(r.vt_cols), 2)
    rec = dict(
        config=name, matrix_seed=12345, call=f"svd_dim_seed(A, {int(dims)}, 12345)", matrix=shape_info,
        diagnostics=dict(non_zero=int(r.non_zero), dimensions=int(r.dimensions), iterations=int(r.iterations),
                         transposed=bool(r.transposed), lanczos_steps=int(r.lanczos_steps),
                         ritz_values_stabilized=int(r.ritz_values_stabilized), significant_values=int(r.significant_values),
                         singular_values=int(r.singular_values), kappa=float(r.kappa), random_seed=int(r.random_seed)),
        d=d, s=[float(x) for x in s],
        alf_head=[float(x) for x in np.ctypeslib.as_array(r.alf, shape=(n,))[:64]],
        bet_head=[float(x) for x in np.ctypeslib.as_array(r.bet, shape=(n + 1,))[:64]],
        ritz_top=[float(x) for x in np.ctypeslib.as_array(r.ritz, shape=(n,))[-8:]],
        counters=dict(n_opb=int(r.n_opb), n_opa=int(r.n_opa), n_purge_fired=int(r.n_purge_fired), n_purge_vecs=int(r.n_purge_vecs),
                      n_restarts=int(r.n_restarts)),
        cpu_seconds=dict(total=r.t_total, opb=r.t_opb, purge=r.t_purge, imtql2=r.t_imtql2, ritvec=r.t_ritvec, wall=wall, gen=gen_s,
                         threads=1, host=os.uname().nodename, nproc=os.cpu_count()),
        sketch=dict(seed=SKETCH_SEED, n_samples=N_SAMPLES, n_proj=N_PROJ,
                    ut_idx=[int(x) for x in iu], vt_idx=[int(x) for x in iv],
                    ut_samples=[[float(x) for x in row] for row in ut[:, iu]],
                    vt_samples=[[float(x) for x in row] for row in vt[:, iv]],
                    ut_proj=[[float(x) for x in row] for row in (ut @ ou.T)],
                    vt_proj=[[float(x) for x in row] for row in (vt @ ov.T)],
                    vt_gram_defect=float(np.abs(vt[:16] @ vt[:16].T - np.eye(min(16, d))).max())),
    )
    L.las2_oracle_free(out)
    path = os.path.join(ROOT, "tests", "golden", f"oracle_{name}.json")
    with open(path, "w") as f:
        json.dump(rec, f)
    print(f"[{name}] oracle {wall:.0f} s: steps={rec['diagnostics']['lanczos_steps']} neig={rec['diagnostics']['ritz_values_stabilized']} "
          f"d={d} nsig={rec['diagnostics']['singular_values']} -> {path}", flush=True)


if __name__ == "__main__":
    for nm in sys.argv[1:]:
        run(nm)
