// testing_abi.cpp -- include/svdlas2_testing.h: host-logic hooks for the CPU test-suite.
#include "../../include/svdlas2_testing.h"

#include <cstring>
#include <vector>

#include "host_math.hpp"
#include "host_pool.hpp"
#include "util.cuh"

using namespace las2;

extern "C" {

void svdlas2_test_start_vector(uint32_t seed, uint64_t n, double *out) {
    ChaCha12Stream g(seed);
    for (uint64_t i = 0; i < n; i++) out[i] = g.next_unit_pm1();
}

double svdlas2_test_hypot(double a, double b) { return hypot_eispack(a, b); }

int svdlas2_test_ql_values(uint64_t n, double *d, double *e, double *bnd) { return ql_values(n, d, e, bnd); }

int svdlas2_test_ql_vectors(uint64_t n, double *d, double *e, double *z, uint64_t *n_rotations, uint64_t *n_sweeps) {
    QlPlan plan;
    int rc = ql_rotations(n, d, e, plan);
    if (n_rotations) *n_rotations = plan.rotations();
    if (n_sweeps) *n_sweeps = plan.sweeps.size();
    if (rc) return rc;
    // CPU replay, same arithmetic as k_ql_replay (tridiag_device.cu): zp holds the physical columns
    std::vector<double> zp(n * n, 0.0);
    for (uint64_t i = 0; i < n; i++) zp[i * n + i] = 1.0;
    for (const QlSweep &sw : plan.sweeps) {
        for (uint64_t row = 0; row < n; row++) {
            double *zr = &zp[row * n];
            uint32_t i = sw.hi;
            double carry = zr[i + 1];
            for (uint32_t t = 0; t < sw.count; t++, i--) {
                const double s = plan.sc[2 * (sw.first + t)], c = plan.sc[2 * (sw.first + t) + 1];
                const double x = zr[i];
                zr[i + 1] = s * x + c * carry;
                carry = c * x - s * carry;
            }
            zr[(uint32_t)(i + 1)] = carry;
        }
    }
    for (uint64_t row = 0; row < n; row++)
        for (uint64_t k = 0; k < n; k++) z[row * n + k] = zp[row * n + plan.column_of[k]];
    return 0;
}

int svdlas2_test_ql_plans_agree(uint64_t n, const double *d, const double *e, uint64_t *n_rotations) {
    // imtqlb with recording against imtql2's own recurrence, on copies of the same (d, e)
    std::vector<double> d1(d, d + n), e1(e, e + n), b1(n, 1.7976931348623157e308), d2(d, d + n), e2(e, e + n);
    std::vector<double> d0(d, d + n), e0(e, e + n), b0(n, 1.7976931348623157e308);
    QlPlan rec, ref;
    const int rc0 = ql_values(n, d0.data(), e0.data(), b0.data());
    const int rc1 = ql_values(n, d1.data(), e1.data(), b1.data(), &rec);
    const int rc2 = ql_rotations(n, d2.data(), e2.data(), ref);
    if (n_rotations) *n_rotations = ref.rotations();
    if (rc0 != rc1) return -1;
    if (rc1 != 0 || rc2 != 0) return (rc1 != 0 && rc2 != 0) ? 1 : -2; // both fail to converge, or they disagree
    // recording must not change what imtqlb returns
    if (std::memcmp(d0.data(), d1.data(), n * sizeof(double)) || std::memcmp(b0.data(), b1.data(), n * sizeof(double))) return -3;
    if (rec.sc.size() != ref.sc.size() || rec.sweeps.size() != ref.sweeps.size() || rec.column_of != ref.column_of) return 0;
    if (!rec.sc.empty() && std::memcmp(rec.sc.data(), ref.sc.data(), rec.sc.size() * sizeof(double))) return 0;
    for (size_t k = 0; k < rec.sweeps.size(); k++)
        if (rec.sweeps[k].hi != ref.sweeps[k].hi || rec.sweeps[k].count != ref.sweeps[k].count || rec.sweeps[k].first != ref.sweeps[k].first)
            return 0;
    return 1;
}

int svdlas2_test_refine_bounds(int *enough, double endl, double endr, double *ritz, double *bnd, uint64_t step,
                               double tol, uint64_t *neig) {
    bool en = *enough != 0;
    size_t ne = 0;
    int rc = refine_bounds(&en, endl, endr, ritz, bnd, step, tol, &ne);
    *enough = en ? 1 : 0;
    if (neig) *neig = ne;
    return rc;
}

int svdlas2_test_host_block(uint64_t bytes, int *node) {
    try {
        if (node) *node = host_pool_node();
        char *p = static_cast<char *>(host_acquire(bytes));
        p[0] = 1;
        if (bytes) p[bytes - 1] = 1;
        host_release(p);
        host_pool_trim();
        return 0;
    } catch (const Error &e) {
        return e.code;
    } catch (...) {
        return SVDLAS2_ERR_ALLOC;
    }
}

void svdlas2_test_cosort(uint64_t n, double *keys, double *vals) { cosort_ascending(n, keys, vals); }

} // extern "C"
