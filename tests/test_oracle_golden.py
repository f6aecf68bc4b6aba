"""The CPU oracle against every golden vector / known-answer test the reference holds for this path
(SURVEY 8(c), App. C).  Runs without a GPU.  The oracle is only trusted because these pass."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
M3 = np.array(GOLDEN["golden_3x3"]["matrix"], dtype=float)


def test_golden_3x3_bit_for_bit(oracle):
    """documented output of svd_dim_seed(&csc, 0, 3141), reference src/lib.rs:259-281 (README.md:235-257)"""
    g = GOLDEN["golden_3x3"]
    r = oracle.svd_dim_seed(sp.csc_matrix(M3), 0, 3141)
    assert r.d == 3
    assert r.s.tolist() == g["s"]            # exact: every one of the 17 printed digits round-trips
    assert r.ut.tolist() == g["ut"]
    assert r.vt.tolist() == g["vt"]
    for k, v in g["diagnostics"].items():
        assert r.diagnostics[k] == v, k


def test_golden_3x3_doctest_assertions(oracle):
    """doctest reference src/lib.rs:218-232"""
    r = oracle.svd_dim_seed(sp.csc_matrix(M3), 0, 3141)
    rec = r.ut.T @ np.diag(r.s) @ r.vt
    assert np.abs(rec - M3).max() < 1e-12
    for got, want in zip(r.s, [123.676578742544, 6.084527896514, 0.287038004183]):
        assert abs(got - want) < 1e-12


@pytest.mark.parametrize("fmt", ["csr", "csc", "coo"])
def test_golden_is_format_independent(oracle, fmt):
    r = oracle.svd_dim_seed(sp.coo_matrix(M3).asformat(fmt), 0, 3141)
    assert r.s.tolist() == GOLDEN["golden_3x3"]["s"]


def test_basic_2x2(oracle):
    """reference src/lib.rs:2183-2214; svd() draws a random seed, the answer does not depend on it"""
    A = sp.coo_matrix(([4.0, 3.0, -5.0], ([0, 1, 1], [0, 0, 1])), shape=(2, 2))
    for seed in (0, 1, 77):
        r = oracle.svd_las2(A, 0, 0, (-1e-30, 1e-30), 1e-6, seed)
        assert r.d == 2 and r.ut.shape == (2, 2) and r.vt.shape == (2, 2)
        assert np.abs(r.ut.T @ np.diag(r.s) @ r.vt - A.toarray()).max() < 1e-12
        assert abs(r.s[0] - 6.3245553203368) < 1e-12 and abs(r.s[1] - 3.1622776601684) < 1e-12


def test_identity_3x3(oracle):
    """reference src/lib.rs:2216-2236"""
    r = oracle.svd(sp.csc_matrix(np.eye(3)))
    assert r.d == 1 and abs(r.s[0] - 1.0) < 1e-12
    assert r.diagnostics["lanczos_steps"] == 2 and r.diagnostics["ritz_values_stabilized"] == 1
    assert r.trace["n_restarts"] == 2


def test_coo_csc_csr_equal(oracle):
    """reference src/lib.rs:2159-2166: the whole Result is == across the three input types"""
    coo = sp.coo_matrix(([3.0, 4.0], ([1, 2], [0, 1])), shape=(4, 4))
    rs = [oracle.svd_dim_seed(coo.asformat(f), 3, 12345) for f in ("coo", "csc", "csr")]
    for r in rs[1:]:
        assert r.d == rs[0].d and np.array_equal(r.ut, rs[0].ut) and np.array_equal(r.s, rs[0].s)
        assert np.array_equal(r.vt, rs[0].vt) and r.diagnostics == rs[0].diagnostics
    assert rs[0].d == 2 and rs[0].diagnostics["lanczos_steps"] == 2
    assert np.allclose(rs[0].s, [4.0, 3.0], atol=1e-14)


def test_wide_input_is_transposed(oracle):
    rng = np.random.default_rng(0)
    A = sp.csr_matrix(rng.standard_normal((2, 5)))
    r = oracle.svd(A)
    assert r.diagnostics["transposed"] and r.ut.shape == (2, 2) and r.vt.shape == (2, 5)
    assert np.allclose(r.s, np.linalg.svd(A.toarray(), compute_uv=False), rtol=1e-14)
    # the orientation test is inclusive: cols >= 1.2 rows (src/lib.rs:606)
    assert oracle.svd_dim_seed(sp.csr_matrix(rng.standard_normal((5, 6))), 0, 1).diagnostics["transposed"]
    assert not oracle.svd_dim_seed(sp.csr_matrix(rng.standard_normal((10, 11))), 0, 1).diagnostics["transposed"]


def test_parameter_clamping(oracle):
    """reference src/lib.rs:583-603"""
    rng = np.random.default_rng(1)
    A = sp.csr_matrix(rng.standard_normal((12, 9)))
    g = oracle.svd_las2(A, 0, 0, (-1e-30, 1e-30), 1e-6, 5).diagnostics
    assert (g["dimensions"], g["iterations"]) == (9, 9)
    g = oracle.svd_las2(A, 4, 2, (-1e-30, 1e-30), 1e-6, 5).diagnostics
    assert (g["dimensions"], g["iterations"]) == (4, 4)          # iterations < dimensions -> dimensions
    g = oracle.svd_las2(A, 100, 100, (-1e-30, 1e-30), -1e-3, 5).diagnostics
    assert (g["dimensions"], g["iterations"]) == (9, 9) and g["kappa"] == 1e-3   # |kappa|
    g = oracle.svd_las2(A, 3, 0, (-1e-30, 1e-30), 0.0, 5).diagnostics
    assert g["kappa"] == np.finfo(float).eps ** 0.75                          # max(|kappa|, eps^(3/4))


def test_error_sites(oracle):
    with pytest.raises(oracle.OracleError) as e:   # src/lib.rs:596-600
        oracle.svd(sp.csr_matrix(np.ones((1, 5))))
    assert e.value.kind == "Las2Error" and "insufficient dimensions: 1" in e.value.message
    with pytest.raises(oracle.OracleError) as e:   # exactly rank 1: assert!(step > 0) in error_bound, :1489
        oracle.svd(sp.csr_matrix(np.ones((3, 3))))
    assert e.value.kind == "Panic"
    with pytest.raises(oracle.OracleError) as e:   # zero matrix: startv, :1127-1129
        oracle.svd_dim_seed(sp.csr_matrix((4, 4)), 2, 7)
    assert e.value.kind == "StartvError"


def test_chacha_known_answers(oracle):
    """published ChaCha20 / ChaCha12 zero-key block 0, and the seed-dependent streams of SURVEY App. B"""
    k = GOLDEN["chacha"]
    assert [f"{w:08x}" for w in oracle.chacha_block([0] * 8, 0, 20)[:8]] == k["chacha20_zero_key_block0_words"]
    b = oracle.chacha_block([0] * 8, 0, 12).astype("<u4").tobytes().hex()
    assert b.startswith(k["chacha12_zero_key_block0_hex_prefix"])
    for seed, first in k["first_values"].items():
        assert oracle.start_vector(int(seed), 4).tolist() == first
    # one u64 per sample, values in [-1, 1), streams of different seeds differ
    v = oracle.start_vector(12345, 10000)
    assert v.min() >= -1.0 and v.max() < 1.0 and abs(v.mean()) < 0.05
    assert not np.array_equal(v[:16], oracle.start_vector(12346, 16))


def test_helpers(oracle):
    assert abs(oracle.pythag(3.0, 4.0) - 5.0) < 2e-15 and oracle.pythag(0.0, 0.0) == 0.0   # iterative EISPACK form
    assert abs(oracle.pythag(1e200, 1e200) - np.sqrt(2) * 1e200) / 1e200 < 1e-15   # no overflow
    a = np.arange(12, dtype=float)
    assert np.array_equal(oracle.rotate(a, 4), np.roll(a, -4))                      # LEFT rotation, :1695-1718
    assert np.array_equal(oracle.rotate(a, 5), np.roll(a, -5))


@pytest.mark.parametrize("n,seed", [(5, 0), (10, 1), (100, 4)])
def test_imtql2_diagonalises(oracle, n, seed):
    rng = np.random.default_rng(seed)
    d, e = rng.standard_normal(n), np.concatenate(([0.0], rng.standard_normal(n - 1)))
    rc, w, z = oracle.imtql2(d, e)
    assert rc == 0
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(T @ z - z * w).max() < 1e-12 and np.abs(z.T @ z - np.eye(n)).max() < 1e-12
    assert np.allclose(w, np.linalg.eigvalsh(T), atol=1e-12)


def test_imtql2_absolute_epsilon_quirk_is_preserved(oracle):
    """compare() is an ABSOLUTE |a-b| < eps test (reference src/lib.rs:810-812) used for QL deflation and for the
    "underflow" exit (:1601, :1635).  On this tridiagonal it deflates early and leaves one eigenvector with a
    5.8e-4 residual; eigenvalues and orthogonality stay at 1e-14.  An independent transliteration of the
    reference gives the same bits, so this is reference behaviour and both oracle and product must keep it."""
    rng = np.random.default_rng(3)
    n = 40
    d, e = rng.standard_normal(n), np.concatenate(([0.0], rng.standard_normal(n - 1)))
    rc, w, z = oracle.imtql2(d, e)
    T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
    res = np.abs(T @ z - z * w).max(axis=0)
    assert rc == 0 and np.argmax(res) == 37 and 5e-4 < res.max() < 7e-4
    assert np.abs(z.T @ z - np.eye(n)).max() < 1e-13 and np.allclose(w, np.linalg.eigvalsh(T), atol=1e-13)


def test_cfg1_like_behaviour_profile(oracle):
    """SURVEY App. C item 6 (indicative numbers from the survey's probe): scaled down 10x to stay fast"""
    from svdlibrs_b200 import synthetic
    A, _ = synthetic.make("cfg1", scale=0.1)
    r = oracle.svd_dim_seed(A, 20, 12345)
    assert r.d == 20 and r.diagnostics["singular_values"] >= 20
    assert r.trace["n_opb"] == r.diagnostics["lanczos_steps"] + 1 + r.d + r.trace["n_restarts"]
    assert r.trace["n_opa"] == r.d
    assert np.abs(r.vt @ r.vt.T - np.eye(20)).max() < 1e-9
    assert np.all(np.diff(r.s) <= 0)
