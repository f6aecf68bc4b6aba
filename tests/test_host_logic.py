"""Host-side logic of the PRODUCT library (svdlibrs_b200/csrc/host_math.cpp) against the oracle, bit for bit.
These pieces steer control flow (step counts, acceptance), so they must round exactly like the reference.
No GPU needed: the hooks of include/svdlas2_testing.h run on the CPU."""
import ctypes as C

import numpy as np
import pytest

f64p = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hooks(las2):
    from svdlibrs_b200 import _lib
    L = _lib.lib()
    L.svdlas2_test_start_vector.argtypes = [C.c_uint32, C.c_uint64, f64p]
    L.svdlas2_test_start_vector.restype = None
    L.svdlas2_test_hypot.argtypes = [C.c_double, C.c_double]
    L.svdlas2_test_hypot.restype = C.c_double
    L.svdlas2_test_ql_values.argtypes = [C.c_uint64, f64p, f64p, f64p]
    L.svdlas2_test_ql_values.restype = C.c_int
    L.svdlas2_test_ql_vectors.argtypes = [C.c_uint64, f64p, f64p, f64p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.svdlas2_test_ql_vectors.restype = C.c_int
    L.svdlas2_test_refine_bounds.argtypes = [C.POINTER(C.c_int), C.c_double, C.c_double, f64p, f64p, C.c_uint64,
                                             C.c_double, C.POINTER(C.c_uint64)]
    L.svdlas2_test_refine_bounds.restype = C.c_int
    L.svdlas2_test_cosort.argtypes = [C.c_uint64, f64p, f64p]
    L.svdlas2_test_cosort.restype = None
    L.svdlas2_test_ql_plans_agree.argtypes = [C.c_uint64, f64p, f64p, C.POINTER(C.c_uint64)]
    L.svdlas2_test_ql_plans_agree.restype = C.c_int
    L.svdlas2_test_host_block.argtypes = [C.c_uint64, C.POINTER(C.c_int)]
    L.svdlas2_test_host_block.restype = C.c_int
    return L


def p(a):
    return a.ctypes.data_as(f64p)


@pytest.mark.parametrize("seed", [0, 1, 3141, 12345, 2**32 - 1])
def test_start_vector_stream_matches_oracle(hooks, oracle, seed):
    n = 1000
    out = np.zeros(n)
    hooks.svdlas2_test_start_vector(seed, n, p(out))
    assert np.array_equal(out, oracle.start_vector(seed, n))


def test_hypot_matches_oracle(hooks, oracle):
    rng = np.random.default_rng(0)
    for a, b in list(rng.standard_normal((200, 2))) + [(0, 0), (1e300, 1e300), (1e-300, 3e-300), (5, 0), (0, -7)]:
        assert hooks.svdlas2_test_hypot(a, b) == oracle.pythag(a, b)


def tridiag(n, seed, clustered=False):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal(n) if not clustered else 1.0 + 1e-9 * rng.standard_normal(n)
    e = np.concatenate(([0.0], np.abs(rng.standard_normal(n - 1)) * (1e-3 if clustered else 1.0)))
    return d, e


@pytest.mark.parametrize("n,seed,clustered", [(2, 0, False), (3, 1, False), (17, 2, False), (64, 3, False),
                                               (64, 4, True), (200, 5, False)])
def test_ql_values_bitwise(hooks, oracle, n, seed, clustered):
    d, e = tridiag(n, seed, clustered)
    bnd = np.full(n, 1.7976931348623157e308)
    rc0, d0, e0, b0 = oracle.imtqlb(d, e, bnd)
    d1, e1, b1 = d.copy(), e.copy(), bnd.copy()
    rc1 = hooks.svdlas2_test_ql_values(n, p(d1), p(e1), p(b1))
    assert rc0 == rc1 == 0
    assert np.array_equal(d0, d1) and np.array_equal(b0, b1)


@pytest.mark.parametrize("n,seed,clustered", [(2, 0, False), (5, 1, False), (33, 2, False), (40, 3, False),
                                               (64, 4, True), (150, 5, False)])
def test_ql_rotation_plan_reproduces_imtql2_bitwise(hooks, oracle, n, seed, clustered):
    """the recorded rotations + column permutation, replayed on an identity, give the reference's z exactly"""
    d, e = tridiag(n, seed, clustered)
    rc0, w0, z0 = oracle.imtql2(d, e)
    d1, e1, z1 = d.copy(), e.copy(), np.zeros((n, n))
    nrot, nsw = C.c_uint64(), C.c_uint64()
    rc1 = hooks.svdlas2_test_ql_vectors(n, p(d1), p(e1), p(z1), C.byref(nrot), C.byref(nsw))
    assert rc0 == rc1 == 0
    assert np.array_equal(w0, d1)
    assert np.array_equal(z0, z1)
    assert nrot.value >= n - 1 and nsw.value >= 1


def test_refine_bounds_bitwise(hooks, oracle):
    """error_bound through a whole oracle run: feed the product's version the same ritz/bnd and compare"""
    rng = np.random.default_rng(9)
    for trial in range(20):
        n = int(rng.integers(2, 60))
        ritz = np.sort(rng.standard_normal(n)) * 10
        if trial % 3 == 0:  # close pairs trigger the merging branches
            ritz[1::4] = ritz[0::4][: len(ritz[1::4])] * (1 + 1e-13)
            ritz = np.sort(ritz)
        bnd = np.abs(rng.standard_normal(n)) * 10.0 ** rng.integers(-18, 0, n)
        tol = 1e-9
        # python restatement of the reference is the oracle's C code; compare against it through a tiny harness
        r1, b1 = ritz.copy(), bnd.copy()
        en = C.c_int(0)
        ne = C.c_uint64()
        rc = hooks.svdlas2_test_refine_bounds(C.byref(en), -1e-30, 1e-30, p(r1), p(b1), n - 1, tol, C.byref(ne))
        assert rc == 0
        want_b, want_ne = _error_bound_py(ritz.copy(), bnd.copy(), n - 1, tol)
        assert np.array_equal(b1, want_b) and ne.value == want_ne
    en = C.c_int(0)
    assert hooks.svdlas2_test_refine_bounds(C.byref(en), -1.0, 1.0, p(np.zeros(1)), p(np.zeros(1)), 0, 0.0, None) == 7


def _error_bound_py(ritz, bnd, step, tol):
    """reference src/lib.rs:1480-1532, in numpy scalars (IEEE double, same operation order)"""
    eps = np.finfo(float).eps
    eps34 = eps ** 0.75
    mid = int(np.argmax(np.abs(bnd[: step + 1])))
    i = ((step + 1) + (step - 1)) // 2
    while i > mid + 1:
        if abs(ritz[i - 1] - ritz[i]) < eps34 * abs(ritz[i]) and bnd[i] > tol and bnd[i - 1] > tol:
            bnd[i - 1] = np.sqrt(bnd[i] * bnd[i] + bnd[i - 1] * bnd[i - 1])
            bnd[i] = 0.0
        i -= 1
    i = ((step + 1) - (step - 1)) // 2
    while i + 1 < mid:
        if abs(ritz[i + 1] - ritz[i]) < eps34 * abs(ritz[i]) and bnd[i] > tol and bnd[i + 1] > tol:
            bnd[i + 1] = np.sqrt(bnd[i] * bnd[i] + bnd[i + 1] * bnd[i + 1])
            bnd[i] = 0.0
        i += 1
    neig = 0
    gapl = ritz[step] - ritz[0]
    for i in range(step + 1):
        gap = gapl
        if i < step:
            gapl = ritz[i + 1] - ritz[i]
        gap = min(gap, gapl)
        if gap > bnd[i]:
            bnd[i] *= bnd[i] / gap
        if bnd[i] <= 16.0 * eps * abs(ritz[i]):
            neig += 1
    return bnd, neig


def test_cosort(hooks):
    rng = np.random.default_rng(2)
    k, v = rng.standard_normal(50), rng.standard_normal(50)
    k[10] = k[20]  # ties keep their order (insertion sort is stable)
    k1, v1 = k.copy(), v.copy()
    hooks.svdlas2_test_cosort(50, p(k1), p(v1))
    order = np.argsort(k, kind="stable")
    assert np.array_equal(k1, k[order]) and np.array_equal(v1, v[order])


@pytest.mark.parametrize("nbytes", [1, 3 << 20, 300 << 20])
def test_host_result_block_without_a_device(hooks, nbytes):
    """The pool of host result buffers (host_pool.cpp) hands out a usable block whether or not a device / its memory
    node can be found: placement on the device's NUMA node is best effort, never a failure."""
    node = C.c_int(-7)
    assert hooks.svdlas2_test_host_block(nbytes, C.byref(node)) == 0
    assert node.value >= -1


@pytest.mark.parametrize("n,seed,clustered", [(1, 0, False), (2, 0, False), (3, 1, False), (17, 2, False), (64, 3, False),
                                               (64, 4, True), (200, 5, False), (500, 6, False), (500, 7, True), (1200, 8, False)])
def test_plan_recorded_by_the_last_analysis_is_imtql2s_plan(hooks, n, seed, clustered):
    """ritvec reuses the rotations the final analysis of T recorded inside imtqlb (src/lib.rs:962-1068) instead of running
    imtql2's recurrence (:1576-1645) again on the same T: the two plans must be bit-identical -- rotations, sweeps and the
    column permutation of the final ordering -- and recording must not change imtqlb's own outputs."""
    d, e = tridiag(n, seed, clustered)
    nrot = C.c_uint64()
    assert hooks.svdlas2_test_ql_plans_agree(n, p(d), p(e), C.byref(nrot)) == 1
    assert nrot.value >= max(0, n - 1)


def test_plan_recording_with_split_and_degenerate_tridiagonals(hooks):
    """zeros on the off-diagonal (T splits inside the iteration), repeated eigenvalues, NaN (both recurrences give up)"""
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(2, 90))
        d = rng.standard_normal(n)
        e = np.concatenate(([0.0], np.abs(rng.standard_normal(n - 1))))
        if trial % 4 == 0:
            e[rng.integers(1, n)] = 0.0
        if trial % 4 == 1:
            d[:] = 1.0
        if trial % 4 == 2:
            e[1:] *= 1e-9
        assert hooks.svdlas2_test_ql_plans_agree(n, p(d), p(e), None) == 1, (trial, n)
    d, e = np.array([1.0, np.nan, 2.0]), np.array([0.0, 0.5, 0.5])
    assert hooks.svdlas2_test_ql_plans_agree(3, p(d), p(e), None) == 1
