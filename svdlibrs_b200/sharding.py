"""Row sharding of the oriented operator for multi-GPU runs: contiguous row ranges of B balanced
by nnz, one range per rank.  Pure host logic (numpy); testable on CPU with gloo."""
from __future__ import annotations

import numpy as np


def partition_rows_by_nnz(indptr, world_size: int, align: int = 1) -> np.ndarray:
    """Boundaries b[0..world_size] with b[0] = 0, b[-1] = rows, such that every range [b[g], b[g+1]) holds about
    nnz / world_size non-zeros (binary search in the row-offset prefix sums).  `align` rounds interior boundaries
    to a multiple (used by generators that build rows in aligned chunks)."""
    indptr = np.asarray(indptr)
    rows = indptr.size - 1
    nnz = int(indptr[-1])
    targets = (np.arange(1, world_size, dtype=np.float64) * nnz / world_size)
    cuts = np.searchsorted(indptr, targets, side="left").astype(np.int64)
    if align > 1:
        cuts = (cuts + align // 2) // align * align
    cuts = np.clip(cuts, 0, rows)
    b = np.concatenate(([0], cuts, [rows])).astype(np.int64)
    return np.maximum.accumulate(b)


def orient(A):
    """The oriented operator as CSR: (B_csr, transposed) with B = A, or A' when cols >= 1.2 rows
    (reference src/lib.rs:606).  CSC input is reinterpreted, never converted, when that is already CSR(B)."""
    import scipy.sparse as sp
    nrows, ncols = A.shape
    transposed = float(ncols) >= float(nrows) * 1.2
    if transposed:
        if A.format == "csc":
            B = sp.csr_matrix((A.data, A.indices, A.indptr), shape=(ncols, nrows))  # CSC(A) arrays == CSR(A')
        else:
            B = A.T.tocsr()
    else:
        B = A.tocsr()
    return B, transposed


def local_shard(B_csr, bounds, rank: int):
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    return B_csr[lo:hi], lo, hi


def broadcast_comm_id(dist, rank: int):
    """rank 0 creates the NCCL id through the library, every rank receives it over the caller's process group."""
    from . import comm_unique_id
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]
