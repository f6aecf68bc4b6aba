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


def powerlaw_csr(nrows: int, ncols: int, nnz: int, seed: int = 12345, alpha_rows: float = 0.6,
                 alpha_cols: float = 0.9, rows: tuple | None = None, chunk_rows: int = 200_000) -> sp.csr_matrix:
    """Rows [lo, hi) of the power-law matrix (all rows when `rows` is None).  Duplicate columns inside a row are
    merged, so the realised nnz is somewhat below the request; the caller logs the actual count."""
    lens = _row_lengths(nrows, nnz, alpha_rows, seed)
    lens = np.minimum(lens, ncols)
    cdf = _column_cdf(ncols, alpha_cols)
    lo, hi = (0, nrows) if rows is None else rows
    indptr = [np.zeros(1, dtype=np.int64)]
    idx_chunks, val_chunks = [], []
    base = 0
    # chunks are aligned to multiples of chunk_rows of the WHOLE matrix and each has its own random stream, so any
    # row range reproduces exactly the rows of the full matrix (a partial chunk is generated whole and trimmed)
    for c0 in range(lo // chunk_rows * chunk_rows, hi, chunk_rows):
        c1 = min(nrows, c0 + chunk_rows)
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
        counts = np.bincount(r, minlength=a1 - a0)
        indptr.append(base + np.cumsum(counts))
        base += int(counts.sum())
        idx_chunks.append(c)
        val_chunks.append(v)
    indptr = np.concatenate(indptr)
    A = sp.csr_matrix((np.concatenate(val_chunks), np.concatenate(idx_chunks), indptr), shape=(hi - lo, ncols))
    A.has_sorted_indices = True
    return A


def make(name: str, seed: int = 12345, scale: float = 1.0, rows: tuple | None = None):
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
        T = powerlaw_csr(ncols, nrows, nnz, seed, rows=rows)
        A = sp.csc_matrix((T.data, T.indices, T.indptr), shape=(nrows, T.shape[0]))
    else:
        A = powerlaw_csr(nrows, ncols, nnz, seed, rows=rows)
    return A, min(dims, min(A.shape))
