// host_math.cpp -- see host_math.hpp.  Build with -ffp-contract=off: the Rust reference never fuses a*b+c.
#include "host_math.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <utility>

#include "../../include/svdlas2.h"

namespace las2 {

double eps34() { return std::pow(kEps, 0.75); } // src/lib.rs:667-669

double hypot_eispack(double a, double b) { // src/lib.rs:873-890
    const double fa = std::fabs(a), fb = std::fabs(b);
    double p = fa > fb ? fa : fb;
    if (!(p > 0.0)) return 0.0;
    const double q = (fa < fb ? fa : fb) / p;
    double r = q * q;
    double t = 4.0 + r;
    while (!near(t, 4.0)) {
        const double s = r / t;
        const double u = 1.0 + 2.0 * s;
        p *= u;
        const double w = s / u;
        r *= w * w;
        t = 4.0 + r;
    }
    return p;
}

size_t argmax_abs(size_t n, const double *x) { // src/lib.rs:847-862
    size_t best = 0;
    for (size_t i = 1; i < n; i++)
        if (std::fabs(x[i]) > std::fabs(x[best])) best = i;
    return best;
}

void cosort_ascending(size_t n, double *keys, double *vals) { // src/lib.rs:815-825
    for (size_t i = 1; i < n; i++)
        for (size_t j = i; j >= 1 && keys[j - 1] > keys[j]; j--) {
            std::swap(keys[j - 1], keys[j]);
            std::swap(vals[j - 1], vals[j]);
        }
}

namespace {

// index of the first negligible sub-diagonal element at or after l (src/lib.rs:980-990, 1595-1605)
inline size_t split_point(size_t l, size_t n, const double *d, const double *e) {
    const size_t last = n - 1;
    size_t m = l;
    while (m < n) {
        if (m == last) break;
        const double test = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (near(test, test + std::fabs(e[m]))) break;
        m++;
    }
    return m;
}

} // namespace

namespace {

// imtql2's final ordering (selection sort, src/lib.rs:1672-1690) applied to the eigenvalues in the order the QL iteration left
// them; the matching column swaps of z become the plan's permutation
void order_columns(size_t n, double *d, QlPlan &plan) {
    for (size_t l = 1; l < n; l++) {
        const size_t i = l - 1;
        size_t k = i;
        double p = d[i];
        for (size_t j = l; j < n; j++)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            std::swap(plan.column_of[i], plan.column_of[k]);
        }
    }
}

} // namespace

int ql_values(size_t n, double *d, double *e, double *bnd, QlPlan *record) { // src/lib.rs:962-1068
    std::vector<double> unsorted; // record: d[l] as the iteration leaves it (imtqlb inserts it among d[0..l) right away)
    if (record) {
        record->sweeps.clear();
        record->sc.clear();
        record->column_of.resize(n);
        for (size_t k = 0; k < n; k++) record->column_of[k] = (uint32_t)k;
        unsorted.assign(d, d + n);
    }
    if (n == 1) return SVDLAS2_OK;
    const size_t last = n - 1;
    bnd[0] = 1.0;
    for (size_t i = 1; i <= last; i++) {
        bnd[i] = 0.0;
        e[i - 1] = e[i];
    }
    e[last] = 0.0;

    for (size_t l = 0; l <= last; l++) {
        for (int iteration = 0;;) {
            const size_t m = split_point(l, n, d, e);
            double p = d[l];
            double f = bnd[l];
            if (m == l) {
                if (record) unsorted[l] = p;
                // d[l] has converged: insert it (and its bound) among the already ordered d[0..l)
                size_t pos = l;
                while (pos >= 1 && p < d[pos - 1]) {
                    d[pos] = d[pos - 1];
                    bnd[pos] = bnd[pos - 1];
                    pos--;
                }
                d[pos] = p;
                bnd[pos] = f;
                break;
            }
            if (iteration == 30) return SVDLAS2_ERR_IMTQLB;
            iteration++;
            // implicit shift
            double g = (d[l + 1] - p) / (2.0 * e[l]);
            double r = hypot_eispack(g, 1.0);
            g = d[m] - p + e[l] / (g + sign_like(r, g));
            double s = 1.0, c = 1.0;
            p = 0.0;
            if (m == 0) return SVDLAS2_ERR_REFERENCE_PANIC; // src/lib.rs:1029
            size_t i = m - 1;
            bool underflow = false;
            QlSweep sw{(uint32_t)i, 0, record ? (uint64_t)(record->sc.size() / 2) : 0};
            for (;;) {
                f = s * e[i];
                const double b = c * e[i];
                r = hypot_eispack(f, g);
                e[i + 1] = r;
                if (near(r, 0.0)) { underflow = true; break; }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                f = bnd[i + 1];
                bnd[i + 1] = s * bnd[i] + c * f;
                bnd[i] = c * bnd[i] - s * f;
                if (record) { // the rotation imtql2 would apply to columns (i, i+1) of z at this point (:1646-1652)
                    record->sc.push_back(s);
                    record->sc.push_back(c);
                    sw.count++;
                }
                if (i == l) break; // covers both `i == 0` and the loop condition `i >= l`
                i--;
            }
            if (record && sw.count) record->sweeps.push_back(sw);
            if (underflow) {
                d[i + 1] -= p;
            } else {
                d[l] -= p;
                e[l] = g;
            }
            e[m] = 0.0;
        }
    }
    if (record) order_columns(n, unsorted.data(), *record);
    return SVDLAS2_OK;
}

int ql_rotations(size_t n, double *d, double *e, QlPlan &plan) { // src/lib.rs:1576-1693 without the z update
    plan.sweeps.clear();
    plan.sc.clear();
    plan.column_of.resize(n);
    for (size_t k = 0; k < n; k++) plan.column_of[k] = (uint32_t)k;
    if (n == 1) return SVDLAS2_OK;
    if (n < 1) return SVDLAS2_ERR_REFERENCE_PANIC;
    const size_t last = n - 1;
    for (size_t i = 1; i < n; i++) e[i - 1] = e[i];
    e[last] = 0.0;
    plan.sc.reserve(std::min<size_t>(4 * n * n, size_t(1) << 26));

    for (size_t l = 0; l < n; l++) {
        for (int iteration = 0;;) {
            const size_t m = split_point(l, n, d, e);
            if (m == l) break;
            if (iteration == 30) return SVDLAS2_ERR_IMTQL2;
            iteration++;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot_eispack(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + sign_like(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            if (m == 0) return SVDLAS2_ERR_REFERENCE_PANIC; // src/lib.rs:1627
            size_t i = m - 1;
            bool underflow = false;
            QlSweep sw{(uint32_t)i, 0, (uint64_t)(plan.sc.size() / 2)};
            for (;;) {
                const double f = s * e[i];
                const double b = c * e[i];
                r = hypot_eispack(f, g);
                e[i + 1] = r;
                if (near(r, 0.0)) { underflow = true; break; }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                plan.sc.push_back(s); // "form vector": rotation of z columns (i, i+1), replayed on the GPU
                plan.sc.push_back(c);
                sw.count++;
                if (i == l) break;
                i--;
            }
            if (sw.count) plan.sweeps.push_back(sw);
            if (underflow) {
                d[i + 1] -= p;
            } else {
                d[l] -= p;
                e[l] = g;
            }
            e[m] = 0.0;
        }
    }
    order_columns(n, d, plan);
    return SVDLAS2_OK;
}

int refine_bounds(bool *enough, double endl, double endr, double *ritz, double *bnd, size_t step, double tol,
                  size_t *neig_out) { // src/lib.rs:1480-1532
    if (step == 0) return SVDLAS2_ERR_REFERENCE_PANIC; // assert!(step > 0) at :1489
    const double e34 = eps34();
    const size_t mid = argmax_abs(step + 1, bnd);

    // merge the bounds of (numerically) multiple Ritz values, from the top down to mid ...
    for (size_t i = ((step + 1) + (step - 1)) / 2; i > mid + 1; i--) {
        if (std::fabs(ritz[i - 1] - ritz[i]) < e34 * std::fabs(ritz[i]) && bnd[i] > tol && bnd[i - 1] > tol) {
            bnd[i - 1] = std::sqrt(bnd[i] * bnd[i] + bnd[i - 1] * bnd[i - 1]);
            bnd[i] = 0.0;
        }
    }
    // ... and from the bottom up to mid
    for (size_t i = ((step + 1) - (step - 1)) / 2; i + 1 < mid; i++) {
        if (std::fabs(ritz[i + 1] - ritz[i]) < e34 * std::fabs(ritz[i]) && bnd[i] > tol && bnd[i + 1] > tol) {
            bnd[i + 1] = std::sqrt(bnd[i] * bnd[i] + bnd[i + 1] * bnd[i + 1]);
            bnd[i] = 0.0;
        }
    }
    // gap refinement and convergence count
    size_t neig = 0;
    double gapl = ritz[step] - ritz[0];
    for (size_t i = 0; i <= step; i++) {
        double gap = gapl;
        if (i < step) gapl = ritz[i + 1] - ritz[i];
        gap = gap < gapl ? gap : gapl; // f64::min
        if (gap > bnd[i]) bnd[i] *= bnd[i] / gap;
        if (bnd[i] <= 16.0 * kEps * std::fabs(ritz[i])) {
            neig++;
            if (!*enough) *enough = endl < ritz[i] && ritz[i] < endr;
        }
    }
    *neig_out = neig;
    return SVDLAS2_OK;
}

// ---------------------------------------------------------------------------------------------- ChaCha12

namespace {
inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
inline void quarter(uint32_t *x, int a, int b, int c, int d) {
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
    x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
    x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
    x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
}
} // namespace

ChaCha12Stream::ChaCha12Stream(uint32_t seed) {
    std::memset(key_, 0, sizeof key_);
    key_[0] = seed; // seed.to_le_bytes() into bytes 0..4 of the 32-byte key (src/lib.rs:1109-1112)
}

void ChaCha12Stream::refill() {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) st[4 + i] = key_[i];
    st[12] = (uint32_t)counter_;
    st[13] = (uint32_t)(counter_ >> 32);
    st[14] = st[15] = 0;
    uint32_t x[16];
    std::memcpy(x, st, sizeof x);
    for (int round = 0; round < 6; round++) { // 6 double rounds = ChaCha12
        quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
        quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) buf_[i] = x[i] + st[i];
    counter_++;
    pos_ = 0;
}

uint64_t ChaCha12Stream::next_u64() {
    if (pos_ >= 16) refill();
    const uint64_t lo = buf_[pos_], hi = buf_[pos_ + 1]; // pos_ stays even: only u64 draws are made
    pos_ += 2;
    return lo | (hi << 32);
}

double ChaCha12Stream::next_unit_pm1() {
    // UniformFloat<f64>::sample_single for -1.0..1.0: 52 random mantissa bits in [1,2), shift to [0,1), scale
    const uint64_t bits = (next_u64() >> 12) | (uint64_t(1023) << 52);
    double v;
    std::memcpy(&v, &bits, sizeof v);
    return (v - 1.0) * 2.0 + -1.0;
}

uint32_t entropy_seed_u32() { // thread_rng().gen::<u32>() (src/lib.rs:580)
    uint32_t v = 0;
    if (FILE *f = std::fopen("/dev/urandom", "rb")) {
        if (std::fread(&v, sizeof v, 1, f) != 1) v = 0;
        std::fclose(f);
    }
    if (v == 0) v = (uint32_t)std::time(nullptr) | 1u;
    return v;
}

} // namespace las2
