"""Text-format loaders: parsing only, no GPU."""
import gzip

import numpy as np
import pytest
import scipy.sparse as sp

from svdlibrs_b200 import io as las2_io


def test_matrix_market_general_symmetric_pattern(tmp_path):
    p = tmp_path / "a.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% comment\n\n3 4 4\n1 1 2.5\n3 4 -1\n2 2 7\n1 1 0.5\n")
    A = las2_io.read_matrix_market(p)
    assert A.format == "coo" and A.shape == (3, 4) and A.nnz == 4  # the duplicate (1,1) stays a separate triplet
    assert np.allclose(A.toarray(), [[3.0, 0, 0, 0], [0, 7, 0, 0], [0, 0, 0, -1]])
    q = tmp_path / "s.mtx"
    q.write_text("%%MatrixMarket matrix coordinate integer symmetric\n3 3 3\n1 1 4\n3 1 2\n3 2 5\n")
    S = las2_io.read_matrix_market(q).toarray()
    assert np.array_equal(S, S.T) and S[2, 0] == 2 and S[0, 2] == 2 and S[0, 0] == 4
    r = tmp_path / "p.mtx.gz"
    with gzip.open(r, "wt") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n2 2 2\n1 2\n2 1\n")
    assert np.array_equal(las2_io.load(r).toarray(), [[0, 1], [1, 0]])
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n")
    with pytest.raises(ValueError):
        las2_io.read_matrix_market(bad)
    arr = tmp_path / "arr.mtx"
    arr.write_text("%%MatrixMarket matrix array real general\n1 1\n1.0\n")
    with pytest.raises(ValueError):
        las2_io.read_matrix_market(arr)


def test_svdlibc_text_formats_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    A = sp.random(7, 5, density=0.4, format="csc", random_state=rng)
    A.sort_indices()
    st = tmp_path / "a.st"
    with open(st, "wt") as f:
        f.write(f"{A.shape[0]} {A.shape[1]} {A.nnz}\n")
        for c in range(A.shape[1]):
            lo, hi = A.indptr[c], A.indptr[c + 1]
            f.write(f"{hi - lo}\n")
            for k in range(lo, hi):
                f.write(f"{A.indices[k]} {float(A.data[k])!r}\n")
    B = las2_io.load(st)
    assert B.format == "csc" and (A != B).nnz == 0
    D = rng.standard_normal((3, 4))
    dt = tmp_path / "d.dt"
    las2_io.write_svdlibc_dense_text(dt, D)
    assert np.array_equal(las2_io.load(dt).toarray(), D)  # 17 significant digits: exact round trip
    s = tmp_path / "s"
    las2_io.write_svdlibc_dense_text(s, np.array([3.0, 2.0, 1.0]))
    assert s.read_text().split() == ["3", "3", "2", "1"]
    with pytest.raises(ValueError):
        las2_io.load(tmp_path / "x.unknown")
