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


def test_bench_workload_shards_tile_the_matrix():
    """bench.py at N > 1: every rank generates only its rows of the oriented operator; together they are the N = 1 matrix
    (tall CSR configs directly, the wide CSC config through its tall transpose)."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name, scale in (("cfg2", 0.01), ("cfg3", 0.005), ("cfg1", 0.2)):
        full = bench.make_workload(name, scale, 0, 1, 1)
        parts = [bench.make_workload(name, scale, r, 3, 1) for r in range(3)]
        assert [p.row_begin for p in parts] == sorted(p.row_begin for p in parts) and parts[0].row_begin == 0
        assert sum(p.local_shape[0] for p in parts) == parts[0].m_global
        assert sum(p.local_nnz for p in parts) == full.local_nnz
        assert all(p.transposed == full.transposed for p in parts)
        if full.fmt in ("csr", "csc") and (name != "cfg1"):
            # power-law configs: the concatenated shards ARE the arrays of the full matrix (CSR of B = the CSC arrays of cfg 3)
            assert np.array_equal(np.concatenate([p.indices for p in parts]), full.indices)
            assert np.array_equal(np.concatenate([p.data for p in parts]), full.data)
