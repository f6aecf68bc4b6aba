"""Sparse-matrix text formats around the hot path (the reference crate itself has no I/O, its
ancestor SVDLIBC's `svd` tool reads these): pure host-side parsing into scipy matrices that `svd*` accept.

  read_matrix_market(path)   coordinate real/integer/pattern, general or symmetric  -> coo_matrix (duplicates kept:
                             CooMatrix semantics, src/lib.rs:2121-2127: every triplet contributes)
  read_svdlibc_sparse_text(path)   SVDLIBC "sparse text" (-r st): header "rows cols nnz", then per COLUMN a count line
                             followed by "row value" lines                                  -> csc_matrix
  read_svdlibc_dense_text(path)    "-r dt": header "rows cols", then rows*cols values       -> csr_matrix
  write_svdlibc_dense_text(path, M)  "-w dt" as the tool writes Ut / Vt / S (rows cols header, one row per line)
  load(path)                 by extension: .mtx/.mm -> Matrix Market, .st -> sparse text, .dt -> dense text
"""
from __future__ import annotations

import gzip
import io as _io
import os

import numpy as np
import scipy.sparse as sp

__all__ = ["read_matrix_market", "read_svdlibc_sparse_text", "read_svdlibc_dense_text", "write_svdlibc_dense_text", "load"]


def _open(path):
    return gzip.open(path, "rt") if str(path).endswith(".gz") else open(path, "rt")


def read_matrix_market(path) -> sp.coo_matrix:
    with _open(path) as f:
        header = f.readline().split()
        if len(header) < 5 or header[0] != "%%MatrixMarket" or header[1].lower() != "matrix":
            raise ValueError(f"{path}: not a Matrix Market file")
        layout, field, symmetry = header[2].lower(), header[3].lower(), header[4].lower()
        if layout != "coordinate":
            raise ValueError(f"{path}: only the coordinate (sparse) layout is supported, got {layout!r}")
        if field not in ("real", "integer", "pattern", "double"):
            raise ValueError(f"{path}: unsupported field {field!r} (real / integer / pattern)")
        if symmetry not in ("general", "symmetric", "skew-symmetric"):
            raise ValueError(f"{path}: unsupported symmetry {symmetry!r}")
        line = f.readline()
        while line and (line.startswith("%") or not line.strip()):
            line = f.readline()
        nrows, ncols, nnz = (int(t) for t in line.split()[:3])
        body = np.loadtxt(f, ndmin=2) if nnz else np.zeros((0, 3))
    if body.shape[0] != nnz:
        raise ValueError(f"{path}: header announces {nnz} entries, file holds {body.shape[0]}")
    rows = body[:, 0].astype(np.int64) - 1
    cols = body[:, 1].astype(np.int64) - 1
    vals = np.ones(nnz) if field == "pattern" else body[:, 2].astype(np.float64)
    if nnz and (rows.min() < 0 or cols.min() < 0 or rows.max() >= nrows or cols.max() >= ncols):
        raise ValueError(f"{path}: index out of range")
    if symmetry != "general":
        off = rows != cols
        sign = -1.0 if symmetry == "skew-symmetric" else 1.0
        rows, cols, vals = (np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]]),
                            np.concatenate([vals, sign * vals[off]]))
    return sp.coo_matrix((vals, (rows, cols)), shape=(nrows, ncols))


def read_svdlibc_sparse_text(path) -> sp.csc_matrix:
    with _open(path) as f:
        tok = f.read().split()
    nrows, ncols, nnz = int(tok[0]), int(tok[1]), int(tok[2])
    indptr = np.zeros(ncols + 1, dtype=np.int64)
    indices = np.empty(nnz, dtype=np.int64)
    data = np.empty(nnz, dtype=np.float64)
    p, k = 3, 0
    for c in range(ncols):
        n = int(tok[p]); p += 1
        for _ in range(n):
            indices[k] = int(tok[p]); data[k] = float(tok[p + 1]); p += 2; k += 1
        indptr[c + 1] = k
    if k != nnz:
        raise ValueError(f"{path}: header announces {nnz} entries, columns hold {k}")
    if nnz and (indices.min() < 0 or indices.max() >= nrows):
        raise ValueError(f"{path}: row index out of range")
    A = sp.csc_matrix((data, indices, indptr), shape=(nrows, ncols))
    A.sort_indices()
    return A


def read_svdlibc_dense_text(path) -> sp.csr_matrix:
    with _open(path) as f:
        tok = f.read().split()
    nrows, ncols = int(tok[0]), int(tok[1])
    vals = np.array(tok[2:2 + nrows * ncols], dtype=np.float64)
    if vals.size != nrows * ncols:
        raise ValueError(f"{path}: expected {nrows * ncols} values, found {vals.size}")
    return sp.csr_matrix(vals.reshape(nrows, ncols))


def write_svdlibc_dense_text(path, M) -> None:
    M = np.atleast_1d(np.asarray(M, dtype=np.float64))
    with open(path, "wt") as f:
        if M.ndim == 1:  # the singular values file: count, then one value per line
            f.write(f"{M.size}\n")
            f.writelines(f"{v:.17g}\n" for v in M)
        else:
            f.write(f"{M.shape[0]} {M.shape[1]}\n")
            for row in M:
                f.write(" ".join(f"{v:.17g}" for v in row) + "\n")


def load(path):
    name = str(path)[:-3] if str(path).endswith(".gz") else str(path)
    ext = os.path.splitext(name)[1].lower()
    if ext in (".mtx", ".mm"):
        return read_matrix_market(path)
    if ext == ".st":
        return read_svdlibc_sparse_text(path)
    if ext == ".dt":
        return read_svdlibc_dense_text(path)
    raise ValueError(f"{path}: unknown extension {ext!r} (.mtx/.mm, .st, .dt, optionally .gz)")
