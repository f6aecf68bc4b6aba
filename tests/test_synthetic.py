"""The synthetic workloads: deterministic in the seed, and any row range reproduces exactly the rows of
the full matrix, so that multi-GPU ranks can each generate only their own shard (scripts/scale_cfg.py, cfg 4)."""
import numpy as np
import scipy.sparse as sp

from svdlibrs_b200 import sharding, synthetic


def test_row_ranges_reproduce_the_full_matrix():
    nrows, ncols, nnz = 50_000, 3_000, 400_000
    full = synthetic.powerlaw_csr(nrows, ncols, nnz, 12345, chunk_rows=7_000)
    lens = np.minimum(synthetic._row_lengths(nrows, nnz, 0.6, 12345), ncols)
    bounds = sharding.partition_rows_by_nnz(np.concatenate([[0], np.cumsum(lens)]), 3)
    parts = [synthetic.powerlaw_csr(nrows, ncols, nnz, 12345, rows=(int(bounds[i]), int(bounds[i + 1])), chunk_rows=7_000)
             for i in range(3)]
    stacked = sp.vstack(parts).tocsr()
    assert stacked.shape == full.shape and stacked.nnz == full.nnz
    assert np.array_equal(stacked.indptr, full.indptr) and np.array_equal(stacked.indices, full.indices)
    assert np.array_equal(stacked.data, full.data)


def test_configs_are_deterministic_and_shaped():
    A, dims = synthetic.make("cfg2", scale=0.01)
    B, _ = synthetic.make("cfg2", scale=0.01)
    assert (A != B).nnz == 0 and A.shape == (10_000, 1_000) and dims == 100
    C3, _ = synthetic.make("cfg3", scale=0.005)
    assert C3.format == "csc" and C3.shape[1] >= 1.2 * C3.shape[0]
    A1, d1 = synthetic.make("cfg1", scale=0.1)
    assert A1.format == "csr" and d1 == 20 and A1.has_sorted_indices


def test_parallel_generation_is_bit_identical():
    """bench.py fans the independent 200k-row chunks over forked workers (cfg 4 is 9.2e8 non-zeros): same bits for any
    worker count and any row range."""
    kw = dict(nrows=60_000, ncols=4_000, nnz=900_000, seed=12345, chunk_rows=7_000)
    a = synthetic.powerlaw_arrays(**kw, workers=1)
    b = synthetic.powerlaw_arrays(**kw, workers=4)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    lo, hi = 12_345, 43_210
    c = synthetic.powerlaw_arrays(**kw, workers=3, rows=(lo, hi))
    assert np.array_equal(c[0], a[0][lo:hi + 1] - a[0][lo])
    assert np.array_equal(c[1], a[1][a[0][lo]:a[0][hi]]) and np.array_equal(c[2], a[2][a[0][lo]:a[0][hi]])


def test_oracle_unit_costs_are_positive(oracle):
    u = oracle.unit_costs(50_000, 4)
    assert all(v > 0 for v in u.values())
