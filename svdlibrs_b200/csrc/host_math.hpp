// host_math.hpp -- the small O(j) / O(j^2) host-side pieces of LAS2 that steer control flow.
//
// north star: "The O(k^2) tridiagonal solvers imtqlb/imtql2 and the Ritz acceptance logic stay on the host."
// They are restated here with EXACTLY the reference's arithmetic (compile with -ffp-contract=off), because
// every branch of the Lanczos driver hangs off their results:
//   near()            compare()        src/lib.rs:810-812   (absolute |b-a| < eps test)
//   hypot_eispack()   svd_pythag()     src/lib.rs:873-890
//   sign_like()       svd_fsign()      src/lib.rs:865-870
//   argmax_abs()      svd_idamax()     src/lib.rs:847-862
//   cosort_ascending  insert_sort()    src/lib.rs:815-825
//   ql_values()       imtqlb()         src/lib.rs:962-1068
//   ql_rotations()    imtql2()         src/lib.rs:1576-1693  -- the d/e recurrences only; the O(js^3) update of
//                                       the eigenvector matrix z is NOT done here: the rotations are recorded and
//                                       replayed on the GPU (tridiag_device.cu), one thread per row of z.
//   refine_bounds()   error_bound()    src/lib.rs:1480-1532
//   ChaCha12Stream    StdRng/gen_range src/lib.rs:1109-1114 (rand 0.8.5 / rand_chacha 0.3, restated)
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace las2 {

constexpr double kEps = 2.220446049250313e-16; // f64::EPSILON

inline bool near(double computed, double expected) {
    double d = expected - computed;
    return (d < 0 ? -d : d) < kEps;
}
double eps34();
double hypot_eispack(double a, double b);
inline double sign_like(double a, double b) { return ((a >= 0.0 && b >= 0.0) || (a < 0.0 && b < 0.0)) ? a : -a; }
size_t argmax_abs(size_t n, const double *x);
void cosort_ascending(size_t n, double *keys, double *vals);

// status: 0 ok, SVDLAS2_ERR_IMTQLB, SVDLAS2_ERR_REFERENCE_PANIC
// record != nullptr: also records the rotation plan imtql2 would produce for the SAME (d, e).  The implicit-QL iterations of
// imtqlb and imtql2 are the same recurrences on the same data (imtqlb merely files every converged eigenvalue among d[0..l)
// at once, positions the iteration never reads again, where imtql2 sorts at the end), so the plan is bit-identical to
// ql_rotations' (tests/test_host_logic.py) and ritvec need not run the O(n^2) chain a second time when T is one block.
struct QlPlan;
int ql_values(size_t n, double *d, double *e, double *bnd, QlPlan *record = nullptr);

// One Givens rotation of columns (i, i+1) of z:  f = z[i+1]; z[i+1] = s*z[i] + c*f; z[i] = c*z[i] - s*f
struct QlSweep {
    uint32_t hi;     // first rotation index applies to columns (hi, hi+1); indices descend by one per rotation
    uint32_t count;  // number of rotations in the sweep
    uint64_t first;  // offset into QlPlan::sc
};
struct QlPlan {
    std::vector<QlSweep> sweeps;
    std::vector<double> sc; // (s, c) pairs, 2 doubles per rotation, in application order
    std::vector<uint32_t> column_of; // after the final ordering: eigenvector k lives in physical column column_of[k]
    uint64_t rotations() const { return sc.size() / 2; }
};
// Runs the implicit-QL iteration of imtql2 on (d, e) (both overwritten: d ends as ascending eigenvalues) and
// records every rotation it would have applied to z, plus the column permutation of the final selection sort.
// status: 0 ok, SVDLAS2_ERR_IMTQL2, SVDLAS2_ERR_REFERENCE_PANIC
int ql_rotations(size_t n, double *d, double *e, QlPlan &plan);

// error_bound(): returns status (0 / SVDLAS2_ERR_REFERENCE_PANIC when step == 0); neig through *neig
int refine_bounds(bool *enough, double endl, double endr, double *ritz, double *bnd, size_t step, double tol,
                  size_t *neig);

// rand 0.8.5 StdRng (ChaCha12, 64-bit counter, stream 0) seeded with a u32 in key bytes 0..4
class ChaCha12Stream {
  public:
    explicit ChaCha12Stream(uint32_t seed);
    uint64_t next_u64();
    double next_unit_pm1(); // gen_range(-1.0..1.0)
  private:
    void refill();
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t buf_[16];
    int pos_ = 16;
};

uint32_t entropy_seed_u32();

} // namespace las2
