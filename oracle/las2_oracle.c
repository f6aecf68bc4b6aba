/*
 * oracle/las2_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into, imported by or executed from the
 * product path; the product fails loudly without its CUDA library instead of falling back to this).
 *
 * Single-threaded CPU restatement, in plain C, of the reference's LAS2 single-vector Lanczos SVD
 * (dfarnham/svdlibrs 0.5.1, /root/reference/src/lib.rs).  Each function cites the reference lines it
 * follows.  Arithmetic order is kept exactly (sequential left-to-right sums, separate multiply and add:
 * build with -ffp-contract=off and without -ffast-math) so that this file reproduces the reference's
 * documented 17-digit output for the 3x3 / seed-3141 example BIT FOR BIT (tests/test_oracle_golden.py).
 *
 * Parity status: PINNED against the reference's own golden vectors and known-answer tests
 * (reference src/lib.rs:172-283, 2159-2236).  The Rust crate itself cannot be built in this image
 * (no cargo/rustc), so there is no oracle/_ref.
 *
 * The start-vector random stream (rand 0.8.5 StdRng) is restated in oracle/chacha12.c.
 */
#include "las2_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[16];
    int pos;
} oracle_rng_t;
void oracle_rng_seed_u32(oracle_rng_t *g, uint32_t seed);
double oracle_rng_unit_pm1(oracle_rng_t *g);

#define MAXLL 2 /* reference src/lib.rs:665 */
#define EPS DBL_EPSILON

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* reference src/lib.rs:667-669 */
static double eps34(void) { return pow(EPS, 0.75); }

/* ---------------------------------------------------------------- sparse operator */

typedef struct {
    int fmt;
    uint64_t nrows, ncols, nnz;
    const uint64_t *a; /* CSR/CSC: major offsets; COO: row indices */
    const uint64_t *b; /* CSR/CSC: minor indices; COO: col indices */
    const double *val;
} smat_t;

/* y = A x (transposed == 0) or y = A' x (transposed != 0); reference src/lib.rs:2041-2063 (CSC),
 * 2077-2099 (CSR), 2113-2129 (COO).  y is zero-filled first, accumulation is in storage order. */
static void svd_opa(const smat_t *A, const double *x, double *y, int transposed)
{
    uint64_t ylen = transposed ? A->ncols : A->nrows;
    uint64_t xlen = transposed ? A->nrows : A->ncols;
    for (uint64_t i = 0; i < ylen; i++) y[i] = 0.0;
    const uint64_t *off = A->a, *idx = A->b;
    const double *v = A->val;
    if (A->fmt == ORACLE_FMT_CSR) {
        if (!transposed) { /* :2086-2091 gather over rows */
            for (uint64_t i = 0; i < ylen; i++) {
                double acc = y[i];
                for (uint64_t j = off[i]; j < off[i + 1]; j++) acc += v[j] * x[idx[j]];
                y[i] = acc;
            }
        } else { /* :2092-2097 scatter */
            for (uint64_t i = 0; i < xlen; i++) {
                double xv = x[i];
                for (uint64_t j = off[i]; j < off[i + 1]; j++) y[idx[j]] += v[j] * xv;
            }
        }
    } else if (A->fmt == ORACLE_FMT_CSC) {
        if (transposed) { /* :2050-2055 gather over columns */
            for (uint64_t i = 0; i < ylen; i++) {
                double acc = y[i];
                for (uint64_t j = off[i]; j < off[i + 1]; j++) acc += v[j] * x[idx[j]];
                y[i] = acc;
            }
        } else { /* :2056-2061 scatter */
            for (uint64_t i = 0; i < xlen; i++) {
                double xv = x[i];
                for (uint64_t j = off[i]; j < off[i + 1]; j++) y[idx[j]] += v[j] * xv;
            }
        }
    } else { /* COO, triplet order; :2120-2128 */
        const uint64_t *ri = A->a, *ci = A->b;
        if (transposed) {
            for (uint64_t k = 0; k < A->nnz; k++) y[ci[k]] += v[k] * x[ri[k]];
        } else {
            for (uint64_t k = 0; k < A->nnz; k++) y[ri[k]] += v[k] * x[ci[k]];
        }
    }
}

/* ---------------------------------------------------------------- state */

typedef struct { /* reference src/lib.rs:707-737 */
    uint64_t n;
    double **vecs;
    uint64_t len, cap;
} store_t;

static void store_push(store_t *s)
{
    if (s->len == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 16;
        s->vecs = (double **)realloc(s->vecs, s->cap * sizeof(double *));
    }
    s->vecs[s->len++] = (double *)calloc(s->n ? s->n : 1, sizeof(double));
}
static void storq(store_t *s, uint64_t idx, const double *v) /* :715-722 */
{
    while (idx + MAXLL >= s->len) store_push(s);
    memcpy(s->vecs[idx + MAXLL], v, s->n * sizeof(double));
}
static void storp(store_t *s, uint64_t idx, const double *v) /* :723-730 */
{
    while (idx >= s->len) store_push(s);
    memcpy(s->vecs[idx], v, s->n * sizeof(double));
}
static const double *retrq(store_t *s, uint64_t idx) { return s->vecs[idx + MAXLL]; } /* :731-733 */
static const double *retrp(store_t *s, uint64_t idx) { return s->vecs[idx]; }         /* :734-736 */
static void store_free(store_t *s)
{
    for (uint64_t i = 0; i < s->len; i++) free(s->vecs[i]);
    free(s->vecs);
}

typedef struct { /* reference src/lib.rs:740-779 */
    uint64_t nrows, ncols;
    int transposed;
    double *w0, *w1, *w2, *w3, *w4, *w5;
    double *alf, *eta, *oldeta, *bet, *bnd, *ritz, *temp;
    /* oracle-only bookkeeping */
    uint64_t n_opb, n_opa, n_purge_fired, n_purge_vecs, n_restarts;
    double t_opb, t_purge;
} work_t;

static double *zeros(uint64_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

static void work_init(work_t *w, uint64_t nrows, uint64_t ncols, int transposed, uint64_t iterations)
{
    memset(w, 0, sizeof *w);
    w->nrows = nrows; w->ncols = ncols; w->transposed = transposed;
    w->w0 = zeros(ncols); w->w1 = zeros(ncols); w->w2 = zeros(ncols);
    w->w3 = zeros(ncols); w->w4 = zeros(ncols); w->w5 = zeros(ncols);
    w->alf = zeros(iterations); w->eta = zeros(iterations); w->oldeta = zeros(iterations);
    w->bet = zeros(1 + iterations); w->ritz = zeros(1 + iterations);
    w->bnd = zeros(1 + iterations);
    for (uint64_t i = 0; i < 1 + iterations; i++) w->bnd[i] = DBL_MAX; /* :775 */
    w->temp = zeros(nrows);
}
static void work_free(work_t *w)
{
    free(w->w0); free(w->w1); free(w->w2); free(w->w3); free(w->w4); free(w->w5);
    free(w->alf); free(w->eta); free(w->oldeta); free(w->bet); free(w->bnd); free(w->ritz); free(w->temp);
}

/* ---------------------------------------------------------------- helpers (reference src/lib.rs:810-924) */

static int compare(double computed, double expected) { return fabs(expected - computed) < EPS; } /* :810-812 */

static void insert_sort(uint64_t n, double *a1, double *a2) /* :815-825 */
{
    for (uint64_t i = 1; i < n; i++) {
        for (uint64_t j = i; j >= 1; j--) {
            if (a1[j - 1] <= a1[j]) break;
            double t = a1[j - 1]; a1[j - 1] = a1[j]; a1[j] = t;
            t = a2[j - 1]; a2[j - 1] = a2[j]; a2[j] = t;
        }
    }
}

static void svd_opb(const smat_t *A, work_t *w, const double *x, double *y) /* :829-837 */
{
    double t0 = now_s();
    svd_opa(A, x, w->temp, w->transposed);
    svd_opa(A, w->temp, y, !w->transposed);
    w->t_opb += now_s() - t0;
    w->n_opb += 1;
}

static void svd_daxpy(uint64_t n, double da, const double *x, double *y) /* :840-844 */
{
    for (uint64_t i = 0; i < n; i++) y[i] += da * x[i];
}

static uint64_t svd_idamax(uint64_t n, const double *x) /* :847-862 */
{
    uint64_t imax = 0;
    for (uint64_t i = 1; i < n; i++)
        if (fabs(x[i]) > fabs(x[imax])) imax = i;
    return imax;
}

static double svd_fsign(double a, double b) /* :865-870 */
{
    return ((a >= 0.0 && b >= 0.0) || (a < 0.0 && b < 0.0)) ? a : -a;
}

static double svd_pythag(double a, double b) /* :873-890 */
{
    double n = fmax(fabs(a), fabs(b));
    if (n > 0.0) {
        double p = n;
        double q = fmin(fabs(a), fabs(b)) / p;
        double r = q * q;
        double t = 4.0 + r;
        while (!compare(t, 4.0)) {
            double s = r / t;
            double u = 1.0 + 2.0 * s;
            p *= u;
            double su = s / u;
            r *= su * su;
            t = 4.0 + r;
        }
        return p;
    }
    return 0.0;
}

static double svd_ddot(uint64_t n, const double *x, const double *y) /* :893-895, strict left-to-right */
{
    double s = 0.0;
    for (uint64_t i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}
static double svd_norm(uint64_t n, const double *x) { return sqrt(svd_ddot(n, x, x)); } /* :898-900 */
static void svd_datx(uint64_t n, double d, const double *x, double *y) /* :903-907 */
{
    for (uint64_t i = 0; i < n; i++) y[i] = d * x[i];
}
static void svd_dscal(uint64_t n, double d, double *x) /* :910-914 */
{
    for (uint64_t i = 0; i < n; i++) x[i] *= d;
}
static void svd_dcopy(uint64_t n, uint64_t offset, const double *x, double *y) /* :917-924, REVERSING copy */
{
    if (n > 0) {
        uint64_t start = n - 1;
        for (uint64_t i = 0; i < n; i++) y[offset + start - i] = x[offset + i];
    }
}

/* ---------------------------------------------------------------- imtqlb (reference src/lib.rs:962-1068) */

static int imtqlb(uint64_t n, double *d, double *e, double *bnd)
{
    if (n == 1) return ORACLE_OK;
    bnd[0] = 1.0;
    uint64_t last = n - 1;
    for (uint64_t i = 1; i <= last; i++) {
        bnd[i] = 0.0;
        e[i - 1] = e[i];
    }
    e[last] = 0.0;

    uint64_t i = 0;
    for (uint64_t l = 0; l <= last; l++) {
        int iteration = 0;
        while (iteration <= 30) {
            uint64_t m = l;
            while (m < n) {
                if (m == last) break;
                double test = fabs(d[m]) + fabs(d[m + 1]);
                if (compare(test, test + fabs(e[m]))) break;
                m += 1;
            }
            double p = d[l];
            double f = bnd[l];
            if (m == l) {
                /* order the eigenvalues */
                int exchange = 1;
                if (l > 0) {
                    i = l;
                    while (i >= 1 && exchange) {
                        if (p < d[i - 1]) {
                            d[i] = d[i - 1];
                            bnd[i] = bnd[i - 1];
                            i -= 1;
                        } else {
                            exchange = 0;
                        }
                    }
                }
                if (exchange) i = 0;
                d[i] = p;
                bnd[i] = f;
                iteration = 31;
            } else {
                if (iteration == 30) return ORACLE_ERR_IMTQLB;
                iteration += 1;
                /* form shift */
                double g = (d[l + 1] - p) / (2.0 * e[l]);
                double r = svd_pythag(g, 1.0);
                g = d[m] - p + e[l] / (g + svd_fsign(r, g));
                double s = 1.0, c = 1.0;
                p = 0.0;
                if (m == 0) return ORACLE_ERR_PANIC; /* :1029 */
                i = m - 1;
                int underflow = 0;
                while (!underflow && i >= l) {
                    f = s * e[i];
                    double b = c * e[i];
                    r = svd_pythag(f, g);
                    e[i + 1] = r;
                    if (compare(r, 0.0)) {
                        underflow = 1;
                        break;
                    }
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
                    if (i == 0) break;
                    i -= 1;
                }
                if (underflow) {
                    d[i + 1] -= p;
                } else {
                    d[l] -= p;
                    e[l] = g;
                }
                e[m] = 0.0;
            }
        }
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- startv / stpone (reference :1098-1147, :1179-1201) */

static int startv(const smat_t *A, work_t *w, uint64_t step, store_t *store, uint32_t random_seed, double *rnm_out)
{
    uint64_t n = w->ncols;
    double rnm2 = svd_ddot(n, w->w0, w->w0);
    for (int id = 0; id < 3; id++) {
        if (id > 0 || step > 0 || compare(rnm2, 0.0)) {
            oracle_rng_t g;
            oracle_rng_seed_u32(&g, random_seed); /* rebuilt from the same seed every time, :1109-1114 */
            for (uint64_t i = 0; i < n; i++) w->w0[i] = oracle_rng_unit_pm1(&g);
        }
        memcpy(w->w3, w->w0, n * sizeof(double));
        /* apply operator to put r in range (essential if m singular) */
        svd_opb(A, w, w->w3, w->w0);
        memcpy(w->w3, w->w0, n * sizeof(double));
        rnm2 = svd_ddot(n, w->w3, w->w3);
        if (rnm2 > 0.0) break;
    }
    if (rnm2 <= 0.0) return ORACLE_ERR_STARTV;

    if (step > 0) {
        for (uint64_t i = 0; i < step; i++) {
            const double *v = retrq(store, i);
            svd_daxpy(n, -svd_ddot(n, w->w3, v), v, w->w0);
        }
        /* make sure q[step] is orthogonal to q[step-1] */
        svd_daxpy(n, -svd_ddot(n, w->w4, w->w0), w->w2, w->w0);
        memcpy(w->w3, w->w0, n * sizeof(double));
        double dot = svd_ddot(n, w->w3, w->w3);
        rnm2 = (dot <= EPS * rnm2) ? 0.0 : dot;
    }
    *rnm_out = sqrt(rnm2);
    return ORACLE_OK;
}

static int stpone(const smat_t *A, work_t *w, store_t *store, uint32_t random_seed, double *rnm_out, double *tol_out)
{
    uint64_t n = w->ncols;
    double rnm;
    int rc = startv(A, w, 0, store, random_seed, &rnm);
    if (rc) return rc;
    if (compare(rnm, 0.0)) return ORACLE_ERR_STPONE;

    /* normalize starting vector */
    svd_datx(n, 1.0 / rnm, w->w0, w->w1);
    svd_dscal(n, 1.0 / rnm, w->w3);

    /* take the first step */
    svd_opb(A, w, w->w3, w->w0);
    w->alf[0] = svd_ddot(n, w->w0, w->w3);
    svd_daxpy(n, -w->alf[0], w->w1, w->w0);
    double t = svd_ddot(n, w->w0, w->w3);
    w->alf[0] += t;
    svd_daxpy(n, -t, w->w1, w->w0);
    memcpy(w->w4, w->w0, n * sizeof(double));
    rnm = svd_norm(n, w->w4);
    double anorm = rnm + fabs(w->alf[0]);
    *rnm_out = rnm;
    *tol_out = sqrt(EPS) * anorm;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- ortbnd / purge (reference :1431-1453, :1356-1400) */

static void ortbnd(work_t *w, uint64_t step, double rnm, double eps1)
{
    if (step < 1) return;
    if (!compare(rnm, 0.0) && step > 1) {
        w->oldeta[0] = (w->bet[1] * w->eta[1] + (w->alf[0] - w->alf[step]) * w->eta[0] - w->bet[step] * w->oldeta[0]) / rnm
                       + eps1;
        if (step > 2) {
            for (uint64_t i = 1; i <= step - 2; i++) {
                w->oldeta[i] = (w->bet[i + 1] * w->eta[i + 1] + (w->alf[i] - w->alf[step]) * w->eta[i]
                                + w->bet[i] * w->eta[i - 1] - w->bet[step] * w->oldeta[i])
                                   / rnm
                               + eps1;
            }
        }
    }
    w->oldeta[step - 1] = eps1;
    double *t = w->oldeta; w->oldeta = w->eta; w->eta = t;
    w->eta[step] = eps1;
}

static void purge(uint64_t n, uint64_t ll, work_t *w, uint64_t step, double *rnm, double tol, store_t *store)
{
    if (step < ll + 2) return;
    double reps = sqrt(EPS);
    double eps1 = EPS * sqrt((double)n);

    /* NB: the scan starts at eta[0], not eta[ll] (reference quirk, :1364) */
    uint64_t k = svd_idamax(step - (ll + 1), w->eta) + ll;
    if (fabs(w->eta[k]) > reps) {
        double t0 = now_s();
        w->n_purge_fired += 1;
        double reps1 = eps1 / reps;
        int iteration = 0;
        int flag = 1;
        while (iteration < 2 && flag) {
            if (*rnm > tol) {
                /* bring in a lanczos vector t and orthogonalize both r and q against it */
                double tq = 0.0, tr = 0.0;
                for (uint64_t i = ll; i < step; i++) {
                    const double *v = retrq(store, i);
                    double t = svd_ddot(n, v, w->w3);
                    tq += fabs(t);
                    svd_daxpy(n, -t, v, w->w1);
                    t = svd_ddot(n, v, w->w4);
                    tr += fabs(t);
                    svd_daxpy(n, -t, v, w->w0);
                    w->n_purge_vecs += 1;
                }
                memcpy(w->w3, w->w1, n * sizeof(double));
                double t = svd_ddot(n, w->w0, w->w3);
                tr += fabs(t);
                svd_daxpy(n, -t, w->w1, w->w0);
                memcpy(w->w4, w->w0, n * sizeof(double));
                *rnm = svd_norm(n, w->w4);
                if (tq <= reps1 && tr <= *rnm * reps1) flag = 0;
            }
            iteration += 1;
        }
        for (uint64_t i = ll; i <= step; i++) {
            w->eta[i] = eps1;
            w->oldeta[i] = eps1;
        }
        w->t_purge += now_s() - t0;
    }
}

/* ---------------------------------------------------------------- lanczos_step (reference :1233-1319) */

static int lanczos_step(const smat_t *A, work_t *w, uint64_t first, uint64_t last, uint64_t *ll, int *enough,
                        double *rnm, double *tol, store_t *store, uint64_t *j_out)
{
    uint64_t n = w->ncols;
    double eps1 = EPS * sqrt((double)n);
    uint64_t j = first;
    double *tp;

    while (j < last) {
        tp = w->w1; w->w1 = w->w2; w->w2 = tp;
        tp = w->w3; w->w3 = w->w4; w->w4 = tp;

        storq(store, j - 1, w->w2);
        if (j - 1 < MAXLL) storp(store, j - 1, w->w4);
        w->bet[j] = *rnm;

        /* restart if invariant subspace is found */
        if (compare(*rnm, 0.0)) {
            w->n_restarts += 1;
            int rc = startv(A, w, j, store, 0, rnm);
            if (rc) return rc;
            if (compare(*rnm, 0.0)) *enough = 1;
        }
        if (*enough) {
            tp = w->w1; w->w1 = w->w2; w->w2 = tp; /* "added by Doug" low-rank fix, :1266-1269 */
            break;
        }

        /* take a lanczos step */
        svd_datx(n, 1.0 / *rnm, w->w0, w->w1);
        svd_dscal(n, 1.0 / *rnm, w->w3);
        svd_opb(A, w, w->w3, w->w0);
        svd_daxpy(n, -*rnm, w->w2, w->w0);
        w->alf[j] = svd_ddot(n, w->w0, w->w3);
        svd_daxpy(n, -w->alf[j], w->w1, w->w0);

        /* orthogonalize against initial lanczos vectors */
        if (j <= MAXLL && fabs(w->alf[j - 1]) > 4.0 * fabs(w->alf[j])) *ll = j;
        uint64_t lim = (j - 1 < *ll) ? j - 1 : *ll;
        for (uint64_t i = 0; i < lim; i++) {
            const double *v1 = retrp(store, i);
            double t = svd_ddot(n, v1, w->w0);
            const double *v2 = retrq(store, i);
            svd_daxpy(n, -t, v2, w->w0);
            w->eta[i] = eps1;
            w->oldeta[i] = eps1;
        }

        /* extended local reorthogonalization */
        double t = svd_ddot(n, w->w0, w->w4);
        svd_daxpy(n, -t, w->w2, w->w0);
        if (w->bet[j] > 0.0) w->bet[j] += t;
        t = svd_ddot(n, w->w0, w->w3);
        svd_daxpy(n, -t, w->w1, w->w0);
        w->alf[j] += t;
        memcpy(w->w4, w->w0, n * sizeof(double));
        *rnm = svd_norm(n, w->w4);
        double anorm = w->bet[j] + fabs(w->alf[j]) + *rnm;
        *tol = sqrt(EPS) * anorm;

        /* update the orthogonality bounds */
        ortbnd(w, j, *rnm, eps1);

        /* restore the orthogonality state when needed */
        purge(w->ncols, *ll, w, j, rnm, *tol, store);
        if (*rnm <= *tol) *rnm = 0.0;
        j += 1;
    }
    *j_out = j;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- error_bound (reference :1480-1532) */

static int error_bound(int *enough, double endl, double endr, double *ritz, double *bnd, uint64_t step, double tol,
                       uint64_t *neig_out)
{
    if (step == 0) return ORACLE_ERR_PANIC; /* assert!(step > 0), :1489 */

    /* massage error bounds for very close ritz values */
    uint64_t mid = svd_idamax(step + 1, bnd);

    uint64_t i = ((step + 1) + (step - 1)) / 2;
    while (i > mid + 1) {
        if (fabs(ritz[i - 1] - ritz[i]) < eps34() * fabs(ritz[i]) && bnd[i] > tol && bnd[i - 1] > tol) {
            bnd[i - 1] = sqrt(bnd[i] * bnd[i] + bnd[i - 1] * bnd[i - 1]);
            bnd[i] = 0.0;
        }
        i -= 1;
    }

    i = ((step + 1) - (step - 1)) / 2;
    while (i + 1 < mid) {
        if (fabs(ritz[i + 1] - ritz[i]) < eps34() * fabs(ritz[i]) && bnd[i] > tol && bnd[i + 1] > tol) {
            bnd[i + 1] = sqrt(bnd[i] * bnd[i] + bnd[i + 1] * bnd[i + 1]);
            bnd[i] = 0.0;
        }
        i += 1;
    }

    /* refine the error bounds */
    uint64_t neig = 0;
    double gapl = ritz[step] - ritz[0];
    for (i = 0; i <= step; i++) {
        double gap = gapl;
        if (i < step) gapl = ritz[i + 1] - ritz[i];
        gap = fmin(gap, gapl);
        if (gap > bnd[i]) bnd[i] *= bnd[i] / gap;
        if (bnd[i] <= 16.0 * EPS * fabs(ritz[i])) {
            neig += 1;
            if (!*enough) *enough = (endl < ritz[i] && ritz[i] < endr);
        }
    }
    *neig_out = neig;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- imtql2 (reference :1576-1693) */

static int imtql2(uint64_t nm, uint64_t n, double *d, double *e, double *z)
{
    if (n == 1) return ORACLE_OK;
    if (n < 1) return ORACLE_ERR_PANIC;
    uint64_t last = n - 1;
    for (uint64_t i = 1; i < n; i++) e[i - 1] = e[i];
    e[last] = 0.0;

    uint64_t nnm = n * nm;
    for (uint64_t l = 0; l < n; l++) {
        int iteration = 0;
        /* look for small sub-diagonal element */
        while (iteration <= 30) {
            uint64_t m = l;
            while (m < n) {
                if (m == last) break;
                double test = fabs(d[m]) + fabs(d[m + 1]);
                if (compare(test, test + fabs(e[m]))) break;
                m += 1;
            }
            if (m == l) break;

            if (iteration == 30) return ORACLE_ERR_IMTQL2;
            iteration += 1;

            /* form shift */
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = svd_pythag(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + svd_fsign(r, g));

            double s = 1.0, c = 1.0, p = 0.0;
            if (m == 0) return ORACLE_ERR_PANIC; /* :1627 */
            uint64_t i = m - 1;
            int underflow = 0;
            while (!underflow && i >= l) {
                double f = s * e[i];
                double b = c * e[i];
                r = svd_pythag(f, g);
                e[i + 1] = r;
                if (compare(r, 0.0)) {
                    underflow = 1;
                } else {
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    /* form vector */
                    for (uint64_t k = 0; k < nnm; k += n) {
                        uint64_t index = k + i;
                        f = z[index + 1];
                        z[index + 1] = s * z[index] + c * f;
                        z[index] = c * z[index] - s * f;
                    }
                    if (i == 0) break;
                    i -= 1;
                }
            }
            /* recover from underflow */
            if (underflow) {
                d[i + 1] -= p;
            } else {
                d[l] -= p;
                e[l] = g;
            }
            e[m] = 0.0;
        }
    }

    /* order the eigenvalues */
    for (uint64_t l = 1; l < n; l++) {
        uint64_t i = l - 1;
        uint64_t k = i;
        double p = d[i];
        for (uint64_t j = l; j < n; j++) {
            if (d[j] < p) {
                k = j;
                p = d[j];
            }
        }
        /* ...and corresponding eigenvectors */
        if (k != i) {
            d[k] = d[i];
            d[i] = p;
            for (uint64_t j = 0; j < nnm; j += n) {
                double t = z[j + i]; z[j + i] = z[j + k]; z[j + k] = t;
            }
        }
    }
    return ORACLE_OK;
}

/* reference :1695-1718 : cycle-leader LEFT rotation by x elements */
static void rotate_array(double *a, uint64_t n, uint64_t x)
{
    uint64_t j = 0, start = 0;
    double t1 = a[0];
    for (uint64_t it = 0; it < n; it++) {
        j = (j >= x) ? j - x : j + n - x;
        double t2 = a[j];
        a[j] = t1;
        if (j == start) {
            j += 1;
            start = j;
            if (j < n) t1 = a[j]; /* the reference reads a[j] here; j == n only after the last move */
        } else {
            t1 = t2;
        }
    }
}

/* ---------------------------------------------------------------- ritvec (reference :1767-1879) */

typedef struct {
    uint64_t d, nsig;
    uint64_t ut_cols, vt_cols;
    double *Ut, *S, *Vt;
} rawrec_t;

static int ritvec(const smat_t *A, uint64_t dimensions, double kappa, work_t *w, uint64_t steps, uint64_t neig,
                  store_t *store, rawrec_t *R, double *t_imtql2)
{
    uint64_t js = steps + 1;
    uint64_t jsq = js * js;
    uint64_t ncols = w->ncols;
    double *s = zeros(jsq);
    for (uint64_t i = 0; i < jsq; i += js + 1) s[i] = 1.0;

    uint64_t vt_len = ncols * dimensions;
    double *Vt = zeros(vt_len);

    svd_dcopy(js, 0, w->alf, Vt);
    svd_dcopy(steps, 1, w->bet, w->w5);

    double t0 = now_s();
    int rc = imtql2(js, js, Vt, w->w5, s);
    *t_imtql2 = now_s() - t0;
    if (rc) {
        free(s); free(Vt);
        return rc;
    }

    uint64_t nsig = 0, x = 0;
    uint64_t id2 = jsq - js;
    for (uint64_t k = 0; k < js; k++) {
        if (w->bnd[k] <= kappa * fabs(w->ritz[k]) && k + 1 > js - neig) {
            x = (x == 0) ? dimensions - 1 : x - 1;
            uint64_t offset = x * ncols;
            for (uint64_t c = 0; c < ncols; c++) Vt[offset + c] = 0.0;
            uint64_t idx = id2 + js;
            for (uint64_t i = 0; i < js; i++) {
                idx -= js;
                if (s[idx] != 0.0) {
                    const double *q = retrq(store, i);
                    double sv = s[idx];
                    for (uint64_t c = 0; c < ncols; c++) Vt[c + offset] += sv * q[c];
                }
            }
            nsig += 1;
        }
        id2 += 1;
    }

    /* rotate the singular vectors and values; x is now the location of the highest singular value */
    if (x > 0) rotate_array(Vt, vt_len, x * ncols);

    uint64_t d = dimensions < nsig ? dimensions : nsig;
    double *S = zeros(d);
    double *Ut = zeros(w->nrows * d);
    /* Vt.value.resize(Vt.cols * d) -- rows beyond d are dropped */

    double *tmp = zeros(ncols);
    for (uint64_t i = 0; i < d; i++) {
        const double *vt_vec = Vt + i * ncols;
        double *ut_vec = Ut + i * w->nrows;

        /* multiply by matrix B first */
        svd_opb(A, w, vt_vec, tmp);
        double t = svd_ddot(ncols, vt_vec, tmp);
        S[i] = sqrt(t);

        svd_daxpy(ncols, -t, vt_vec, tmp);
        w->bnd[js] = svd_norm(ncols, tmp) * (1.0 / S[i]);

        /* multiply by matrix A to get (scaled) left s-vector */
        svd_opa(A, vt_vec, ut_vec, w->transposed);
        w->n_opa += 1;
        svd_dscal(w->nrows, 1.0 / S[i], ut_vec);
    }
    free(tmp);
    free(s);

    R->d = d; R->nsig = nsig;
    R->Ut = Ut; R->ut_cols = w->nrows;
    R->S = S;
    R->Vt = (double *)realloc(Vt, ((d * ncols) > 0 ? d * ncols : 1) * sizeof(double));
    R->vt_cols = ncols;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- lanso (reference :1925-2017) */

static int lanso(const smat_t *A, uint64_t dim, uint64_t iterations, const double end_interval[2], work_t *w,
                 uint64_t *neig, store_t *store, uint32_t random_seed, uint64_t *steps_out)
{
    double endl = end_interval[0], endr = end_interval[1];

    /* take the first step */
    double rnm, tol;
    int rc = stpone(A, w, store, random_seed, &rnm, &tol);
    if (rc) return rc;

    double eps1 = EPS * sqrt((double)w->ncols);
    w->eta[0] = eps1;
    w->oldeta[0] = eps1;
    uint64_t ll = 0;
    uint64_t first = 1;
    uint64_t last = (dim > 8 ? dim : 8) + dim;
    if (iterations < last) last = iterations;
    int enough = 0;
    uint64_t j = 0;
    uint64_t intro = 0;

    while (!enough) {
        if (rnm <= tol) rnm = 0.0;

        /* the actual lanczos loop */
        uint64_t steps;
        rc = lanczos_step(A, w, first, last, &ll, &enough, &rnm, &tol, store, &steps);
        if (rc) return rc;
        j = enough ? steps - 1 : last - 1;

        first = j + 1;
        w->bet[first] = rnm;

        /* analyze T */
        uint64_t l = 0;
        for (uint64_t id2 = 0; id2 < j; id2++) {
            if (l > j) break;

            uint64_t i = l;
            while (i <= j) {
                if (compare(w->bet[i + 1], 0.0)) break;
                i += 1;
            }
            if (i > j) i = j;

            /* now i is at the end of an unreduced submatrix */
            uint64_t sz = i - l;
            svd_dcopy(sz + 1, l, w->alf, w->ritz);
            svd_dcopy(sz, l + 1, w->bet, w->w5);

            rc = imtqlb(sz + 1, w->ritz + l, w->w5 + l, w->bnd + l);
            if (rc) return rc;

            for (uint64_t m = l; m <= i; m++) w->bnd[m] = rnm * fabs(w->bnd[m]);
            l = i + 1;
        }

        /* sort eigenvalues into increasing order */
        insert_sort(j + 1, w->ritz, w->bnd);

        rc = error_bound(&enough, endl, endr, w->ritz, w->bnd, j, tol, neig);
        if (rc) return rc;

        /* should we stop? */
        if (*neig < dim) {
            if (*neig == 0) {
                last = first + 9;
                intro = first;
            } else {
                uint64_t ext = 1 + ((j - intro) * (dim - *neig)) / *neig;
                last = first + (ext > 3 ? ext : 3);
            }
            if (iterations < last) last = iterations;
        } else {
            enough = 1;
        }
        enough = enough || first >= iterations;
    }
    storq(store, j, w->w1);
    *steps_out = j;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- svdLAS2 driver (reference :570-655) */

static uint32_t entropy_u32(void)
{
    uint32_t v = 0;
    FILE *f = fopen("/dev/urandom", "rb");
    if (f) {
        if (fread(&v, sizeof v, 1, f) != 1) v = 0;
        fclose(f);
    }
    if (!v) v = (uint32_t)time(NULL) | 1u;
    return v;
}

int las2_oracle_svd(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                    const double *values, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                    double kappa, uint32_t random_seed, oracle_result_t **out)
{
    double t_begin = now_s();
    oracle_result_t *res = (oracle_result_t *)calloc(1, sizeof *res);
    *out = res;
    smat_t A = {fmt, nrows, ncols, nnz, a, b, values};

    if (!(random_seed > 0)) random_seed = entropy_u32(); /* thread_rng().gen(), :578-581 */

    uint64_t mn = nrows < ncols ? nrows : ncols;
    if (dimensions == 0 || dimensions > mn) dimensions = mn;
    if (iterations == 0 || iterations > mn) iterations = mn;
    else if (iterations < dimensions) iterations = dimensions;

    res->non_zero = nnz; res->dimensions = dimensions; res->iterations = iterations;
    res->end_interval[0] = end_interval[0]; res->end_interval[1] = end_interval[1];
    res->random_seed = random_seed;

    if (dimensions < 2) {
        snprintf(res->message, sizeof res->message, "svdLAS2: insufficient dimensions: %llu",
                 (unsigned long long)dimensions);
        return ORACLE_ERR_LAS2;
    }

    /* If the matrix is wide, the SVD is computed on its transpose for speed (:606-608) */
    int transposed = (double)ncols >= ((double)nrows * 1.2);
    uint64_t w_nrows = transposed ? ncols : nrows;
    uint64_t w_ncols = transposed ? nrows : ncols;
    res->transposed = transposed;

    work_t w;
    work_init(&w, w_nrows, w_ncols, transposed, iterations);
    store_t store = {w_ncols, NULL, 0, 0};

    uint64_t neig = 0, steps = 0;
    int rc = lanso(&A, dimensions, iterations, end_interval, &w, &neig, &store, random_seed, &steps);
    rawrec_t R;
    memset(&R, 0, sizeof R);
    double t_imtql2 = 0.0, t_rv0 = now_s();
    if (!rc) {
        kappa = fmax(fabs(kappa), eps34()); /* :627 */
        rc = ritvec(&A, dimensions, kappa, &w, steps, neig, &store, &R, &t_imtql2);
    }
    res->t_ritvec = now_s() - t_rv0;
    if (rc) {
        static const char *names[] = {"ok", "imtqlb no convergence to an eigenvalue after 30 iterations",
                                      "rnm2 <= 0.0", "rnm == 0.0",
                                      "imtql2 no convergence to an eigenvalue after 30 iterations", "las2", "shape",
                                      "reference assert! would panic here"};
        snprintf(res->message, sizeof res->message, "%s", names[rc <= 7 ? rc : 7]);
        work_free(&w);
        store_free(&store);
        return rc;
    }

    if (transposed) { /* :631-633 */
        double *t = R.Ut; R.Ut = R.Vt; R.Vt = t;
        uint64_t c = R.ut_cols; R.ut_cols = R.vt_cols; R.vt_cols = c;
    }
    res->d = R.d;
    res->ut = R.Ut; res->ut_cols = R.ut_cols;
    res->s = R.S;
    res->vt = R.Vt; res->vt_cols = R.vt_cols;
    res->lanczos_steps = steps + 1;
    res->ritz_values_stabilized = neig;
    res->significant_values = R.d;  /* sic: :648 */
    res->singular_values = R.nsig;  /* sic: :649 */
    res->kappa = kappa;

    uint64_t js = steps + 1;
    res->trace_len = js;
    res->alf = zeros(js); res->bet = zeros(js + 1); res->ritz = zeros(js); res->bnd = zeros(js);
    memcpy(res->alf, w.alf, js * sizeof(double));
    memcpy(res->bet, w.bet, (js + 1) * sizeof(double)); /* bet has iterations+1 >= js+1 entries */
    memcpy(res->ritz, w.ritz, js * sizeof(double));
    memcpy(res->bnd, w.bnd, js * sizeof(double));
    res->n_opb = w.n_opb; res->n_opa = w.n_opa;
    res->n_purge_fired = w.n_purge_fired; res->n_purge_vecs = w.n_purge_vecs; res->n_restarts = w.n_restarts;
    res->t_opb = w.t_opb; res->t_purge = w.t_purge; res->t_imtql2 = t_imtql2;
    work_free(&w);
    store_free(&store);
    res->t_total = now_s() - t_begin;
    return ORACLE_OK;
}

void las2_oracle_free(oracle_result_t *r)
{
    if (!r) return;
    free(r->ut); free(r->s); free(r->vt);
    free(r->alf); free(r->bet); free(r->ritz); free(r->bnd);
    free(r);
}

/* ---------------------------------------------------------------- building blocks for kernel parity tests */

void las2_oracle_opa(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                     const double *values, const double *x, double *y, int transposed)
{
    smat_t A = {fmt, nrows, ncols, nnz, a, b, values};
    svd_opa(&A, x, y, transposed);
}

void las2_oracle_opb(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                     const double *values, const double *x, double *y, double *temp, int transposed)
{
    smat_t A = {fmt, nrows, ncols, nnz, a, b, values};
    svd_opa(&A, x, temp, transposed);
    svd_opa(&A, temp, y, !transposed);
}

/* Unit costs of the O(N) loops of the path on this host, for bench.py's BOUNDED CPU sample of configs whose full
 * solve takes hours (time a few svd_opb, one purge pass, imtql2 at the final js, extrapolate with the
 * counts of the full solve).  The loops are the oracle's own (same helpers, same access pattern: `nvec` stored
 * vectors of length n streamed once each).  out[0] = seconds per stored vector visited by purge (two ddot + two
 * daxpy, reference :1374-1382); out[1] = seconds of the BLAS-1 chain of one lanczos_step incl. the storq copy
 * (:1248-1304); out[2] = seconds per (stored vector, accepted Ritz vector) term of the contraction (:1813-1815). */
void las2_oracle_unit_costs(uint64_t n, uint64_t nvec, double out[3])
{
    if (nvec < 1) nvec = 1;
    double **q = (double **)malloc(nvec * sizeof(double *));
    double *w[6];
    uint64_t state = 0x9E3779B97F4A7C15ull;
    for (uint64_t v = 0; v < nvec + 6; v++) {
        double *p = (double *)malloc((n ? n : 1) * sizeof(double));
        for (uint64_t i = 0; i < n; i++) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            p[i] = ((double)(state >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 1e-3;
        }
        if (v < nvec) q[v] = p; else w[v - nvec] = p;
    }
    double *w0 = w[0], *w1 = w[1], *w2 = w[2], *w3 = w[3], *w4 = w[4], *vt = w[5];
    volatile double sink = 0.0;

    double t0 = now_s();
    for (uint64_t i = 0; i < nvec; i++) { /* purge, :1374-1382 */
        double t = svd_ddot(n, q[i], w3);
        svd_daxpy(n, -t, q[i], w1);
        t = svd_ddot(n, q[i], w4);
        svd_daxpy(n, -t, q[i], w0);
        sink += t;
    }
    out[0] = (now_s() - t0) / (double)nvec;

    t0 = now_s();
    {   /* one lanczos_step without its svd_opb and purge, :1248-1304 */
        memcpy(q[0], w2, n * sizeof(double)); /* storq */
        svd_datx(n, 1.0 / 3.0, w0, w1);
        svd_dscal(n, 1.0 / 3.0, w3);
        svd_daxpy(n, -1e-3, w2, w0);
        double a = svd_ddot(n, w0, w3);
        svd_daxpy(n, -a, w1, w0);
        double t = svd_ddot(n, w0, w4);
        svd_daxpy(n, -t, w2, w0);
        t = svd_ddot(n, w0, w3);
        svd_daxpy(n, -t, w1, w0);
        memcpy(w4, w0, n * sizeof(double));
        sink += svd_norm(n, w4) + a;
    }
    out[1] = now_s() - t0;

    t0 = now_s();
    for (uint64_t i = 0; i < nvec; i++) { /* ritvec contraction, :1813-1815 */
        const double *qq = q[i];
        double sv = 1e-3 * (double)(i + 1);
        for (uint64_t c = 0; c < n; c++) vt[c] += sv * qq[c];
    }
    out[2] = (now_s() - t0) / (double)nvec;
    sink += vt[n ? n - 1 : 0];
    (void)sink;
    for (uint64_t v = 0; v < nvec; v++) free(q[v]);
    for (int k = 0; k < 6; k++) free(w[k]);
    free(q);
}

int las2_oracle_imtqlb(uint64_t n, double *d, double *e, double *bnd) { return imtqlb(n, d, e, bnd); }
int las2_oracle_imtql2(uint64_t nm, uint64_t n, double *d, double *e, double *z) { return imtql2(nm, n, d, e, z); }
double las2_oracle_pythag(double a, double b) { return svd_pythag(a, b); }
void las2_oracle_rotate(double *a, uint64_t len, uint64_t x) { rotate_array(a, len, x); }
