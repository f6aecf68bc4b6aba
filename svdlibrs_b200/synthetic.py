"""Synthetic workloads of BASELINE.json's shapes.  Pure numpy/scipy, deterministic in `seed`,
generated on the host because the reference-facing call takes HOST buffers (nalgebra-sparse matrices).

  uniform_csr      cfg 1: uniform-random CSR, values U(0,1)
  powerlaw_csr     cfg 2/4/5: LSA-shaped: row lengths ~ Zipf(alpha_rows), column popularity ~ Zipf(alpha_cols)
                   with popular columns at low indices (term ids sorted by frequency, as LSA vocabularies are),
                   values ln(1 + geometric(1/2)) (tf-like, positive)
  wide_csc         cfg 3: the same generator emitted as CSC of a wide matrix (exercises the transpose path)

`rows=(lo, hi)` generates only that row range of the SAME matrix (multi-GPU ranks build their own shard).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: (generator, kwargs, dims)
    "cfg1": dict(kind="uniform", nrows=50_000, ncols=10_000, nnz=500_000, dims=20, fmt="csr"),
    "cfg2": dict(kind="powerlaw", nrows=1_000_000, ncols=100_000, nnz=50_000_000, dims=100, fmt="csr"),
    "cfg3": dict(kind="powerlaw", nrows=200_000, ncols=2_000_000, nnz=100_000_000, dims=100, fmt="csc"),
    "cfg4": dict(kind="powerlaw", nrows=10_000_000, ncols=2_000_000, nnz=1_000_000_000, dims=200, fmt="csr"),
    "cfg5": dict(kind="powerlaw", nrows=2_000_000, ncols=500_000, nnz=200_000_000, dims=1000, fmt="csr"),
}


def uniform_csr(nrows: int, ncols: int, nnz: int, seed: int = 12345) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    per_row = np.full(nrows, nnz // nrows, dtype=np.int64)
    per_row[: nnz - per_row.sum()] += 1
    rows = np.repeat(np.arange(nrows, dtype=np.int64), per_row)
    cols = rng.integers(0, ncols, size=rows.size, dtype=np.int64)
    vals = rng.random(rows.size)
    keys = np.unique(rows * ncols + cols, return_index=True)
    rows, cols, vals = keys[0] // ncols, keys[0] % ncols, vals[keys[1]]
    A = sp.csr_matrix((vals, (rows, cols)), shape=(nrows, ncols))
    A.sort_indices()
    return A


def _row_lengths(nrows: int, nnz: int, alpha: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, 1])
    ranks = rng.permutation(nrows).astype(np.float64) + 1.0
    w = ranks ** (-alpha)
    lens = np.maximum(1, np.rint(w * (nnz / w.sum()))).astype(np.int64)
    return lens


def _column_cdf(ncols: int, alpha: float) -> np.ndarray:
    w = (np.arange(ncols, dtype=np.float64) + 1.0) ** (-alpha)
    cdf = np.cumsum(w)
    return cdf / cdf[-1]


def _gen_chunk(c0: int, c1: int, lo: int, hi: int, lens: np.ndarray, cdf: np.ndarray, ncols: int, seed: int):
    """Rows [max(lo, c0), min(hi, c1)) of the matrix: (per-row counts, column indices, values).  A chunk is aligned to
    chunk_rows of the WHOLE matrix and has its own random stream, so any row range reproduces exactly the rows of the full
    matrix (a partial chunk is generated whole and trimmed) and chunks can be generated in any order / process."""
    rng = np.random.default_rng([seed, 2, c0])
    ln = lens[c0:c1]
    r = np.repeat(np.arange(c1 - c0, dtype=np.int64), ln)
    c = np.searchsorted(cdf, rng.random(r.size), side="right").astype(np.int64)
    np.minimum(c, ncols - 1, out=c)
    v = np.log1p(rng.geometric(0.5, size=r.size).astype(np.float64))
    key, first = np.unique(r * ncols + c, return_index=True)
    r, c, v = key // ncols, key % ncols, v[first]
    a0, a1 = max(lo, c0) - c0, min(hi, c1) - c0  # rows of this chunk that belong to [lo, hi)
    if a0 > 0 or a1 < c1 - c0:
        keep = (r >= a0) & (r < a1)
        r, c, v = r[keep] - a0, c[keep], v[keep]
    return np.bincount(r, minlength=a1 - a0), c, v


_PAR = {}  # state inherited by forked generator workers (shared anonymous mappings + read-only tables)


def _par_task(k: int) -> int:
    P = _PAR
    c0, c1 = P["chunks"][k]
    counts, c, v = _gen_chunk(c0, c1, P["lo"], P["hi"], P["lens"], P["cdf"], P["ncols"], P["seed"])
    o, n = int(P["ub_off"][k]), int(c.size)
    P["idx"][o:o + n] = c
    P["val"][o:o + n] = v
    r0 = max(P["lo"], c0) - P["lo"]
    P["cnt"][r0:r0 + counts.size] = counts
    return n


def _shared_array(n: int, dtype) -> np.ndarray:
    """Anonymous MAP_SHARED mapping viewed as a numpy array: forked workers write into it, the parent reads the result
    (no pickling of the 15 GB of cfg 4, no dependency on the size of /dev/shm)."""
    import mmap
    nbytes = max(1, int(n)) * np.dtype(dtype).itemsize
    mm = mmap.mmap(-1, nbytes)
    return np.frombuffer(mm, dtype=dtype, count=int(n))


def powerlaw_arrays(nrows: int, ncols: int, nnz: int, seed: int = 12345, alpha_rows: float = 0.6,
                    alpha_cols: float = 0.9, rows: tuple | None = None, chunk_rows: int = 200_000, workers: int = 1):
    """CSR arrays (indptr int64, indices int64, data float64) of rows [lo, hi) of the power-law matrix.  `workers` > 1
    generates the independent 200k-row chunks in forked processes; the result is bit-identical for any worker count.
    Must be called before the process initialises CUDA when workers > 1 (fork)."""
    lens = _row_lengths(nrows, nnz, alpha_rows, seed)
    lens = np.minimum(lens, ncols)
    cdf = _column_cdf(ncols, alpha_cols)
    lo, hi = (0, nrows) if rows is None else rows
    chunks = [(c0, min(nrows, c0 + chunk_rows)) for c0 in range(lo // chunk_rows * chunk_rows, hi, chunk_rows)]
    if workers <= 1 or len(chunks) <= 1:
        cnts, idx_chunks, val_chunks = [], [], []
        for c0, c1 in chunks:
            counts, c, v = _gen_chunk(c0, c1, lo, hi, lens, cdf, ncols, seed)
            cnts.append(counts); idx_chunks.append(c); val_chunks.append(v)
        counts = np.concatenate(cnts) if cnts else np.zeros(0, dtype=np.int64)
        indptr = np.concatenate([np.zeros(1, dtype=np.int64), np.cumsum(counts)])
        idx = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, dtype=np.int64)
        val = np.concatenate(val_chunks) if val_chunks else np.zeros(0)
        return indptr, idx, val
    import multiprocessing as mp
    # upper bound of a chunk's entries: the requested lengths of its kept rows (duplicates only remove entries)
    ub = np.array([int(lens[max(lo, c0):min(hi, c1)].sum()) for c0, c1 in chunks], dtype=np.int64)
    ub_off = np.concatenate([[0], np.cumsum(ub)])
    idx, val, cnt = _shared_array(ub_off[-1], np.int64), _shared_array(ub_off[-1], np.float64), _shared_array(hi - lo, np.int64)
    _PAR.update(chunks=chunks, lo=lo, hi=hi, lens=lens, cdf=cdf, ncols=ncols, seed=seed, ub_off=ub_off, idx=idx, val=val, cnt=cnt)
    try:
        with mp.get_context("fork").Pool(min(workers, len(chunks))) as pool:
            got = pool.map(_par_task, range(len(chunks)), chunksize=1)
    finally:
        _PAR.clear()
    # compact in place (destination never ahead of the source)
    pos = 0
    for k, n in enumerate(got):
        s0 = int(ub_off[k])
        if s0 != pos and n:
            idx[pos:pos + n] = idx[s0:s0 + n].copy() if s0 < pos + n else idx[s0:s0 + n]
            val[pos:pos + n] = val[s0:s0 + n].copy() if s0 < pos + n else val[s0:s0 + n]
        pos += n
    indptr = np.concatenate([np.zeros(1, dtype=np.int64), np.cumsum(cnt)])
    assert int(indptr[-1]) == pos
    return indptr, idx[:pos], val[:pos]


def powerlaw_csr(nrows: int, ncols: int, nnz: int, seed: int = 12345, alpha_rows: float = 0.6,
                 alpha_cols: float = 0.9, rows: tuple | None = None, chunk_rows: int = 200_000, workers: int = 1) -> sp.csr_matrix:
    """Rows [lo, hi) of the power-law matrix (all rows when `rows` is None).  Duplicate columns inside a row are
    merged, so the realised nnz is somewhat below the request; the caller logs the actual count."""
    indptr, idx, val = powerlaw_arrays(nrows, ncols, nnz, seed, alpha_rows, alpha_cols, rows, chunk_rows, workers)
    lo, hi = (0, nrows) if rows is None else rows
    A = sp.csr_matrix((val, idx, indptr), shape=(hi - lo, ncols))
    A.has_sorted_indices = True
    return A


def config_dims(name: str, scale: float = 1.0):
    """(nrows, ncols, requested nnz, dims) of a named config at `scale`."""
    cfg = CONFIGS[name]
    return (max(8, int(cfg["nrows"] * scale)), max(8, int(cfg["ncols"] * scale)), max(16, int(cfg["nnz"] * scale)), cfg["dims"])


def requested_row_offsets(name: str, scale: float = 1.0, seed: int = 12345) -> np.ndarray:
    """Prefix sums of the REQUESTED row lengths of a tall power-law config (the realised ones differ by the merged
    duplicate columns, a few per cent): what multi-GPU ranks balance their row ranges by without generating the matrix."""
    nrows, ncols, nnz, _ = config_dims(name, scale)
    lens = np.minimum(_row_lengths(nrows, nnz, 0.6, seed), ncols)
    return np.concatenate([[0], np.cumsum(lens)])


def make(name: str, seed: int = 12345, scale: float = 1.0, rows: tuple | None = None, workers: int = 1):
    """Matrix of a named BASELINE config (optionally scaled down in every dimension for tests). Returns (A, dims)."""
    cfg = CONFIGS[name]
    nrows = max(8, int(cfg["nrows"] * scale))
    ncols = max(8, int(cfg["ncols"] * scale))
    nnz = max(16, int(cfg["nnz"] * scale))
    dims = cfg["dims"]
    if cfg["kind"] == "uniform":
        A = uniform_csr(nrows, ncols, nnz, seed)
    elif cfg["fmt"] == "csc":
        # wide CSC: generate the tall transpose as CSR and reinterpret its arrays as CSC of the wide matrix
        T = powerlaw_csr(ncols, nrows, nnz, seed, rows=rows, workers=workers)
        A = sp.csc_matrix((T.data, T.indices, T.indptr), shape=(nrows, T.shape[0]))
    else:
        A = powerlaw_csr(nrows, ncols, nnz, seed, rows=rows, workers=workers)
    return A, min(dims, min(A.shape))
