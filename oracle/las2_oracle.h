/*
 * oracle/las2_oracle.h -- TEST INFRASTRUCTURE ONLY.
 * C interface of the CPU oracle (single-threaded restatement of the reference's LAS2 path,
 * reference src/lib.rs:570-2130).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never does.
 */
#ifndef LAS2_ORACLE_H
#define LAS2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_FMT_CSR = 0, ORACLE_FMT_CSC = 1, ORACLE_FMT_COO = 2 };

/* status codes (mirror SvdLibError, reference src/error.rs:8-26, plus the reference's panics) */
enum {
    ORACLE_OK = 0,
    ORACLE_ERR_IMTQLB = 1,
    ORACLE_ERR_STARTV = 2,
    ORACLE_ERR_STPONE = 3,
    ORACLE_ERR_IMTQL2 = 4,
    ORACLE_ERR_LAS2 = 5,
    ORACLE_ERR_PANIC = 7 /* an assert! of the reference would have fired (e.g. src/lib.rs:1489) */
};

typedef struct {
    /* SvdRec (reference src/lib.rs:455-461) */
    uint64_t d;
    uint64_t ut_cols; /* rows(A) */
    uint64_t vt_cols; /* cols(A) */
    double *ut;       /* d x ut_cols, row-major */
    double *s;        /* d */
    double *vt;       /* d x vt_cols, row-major */
    /* Diagnostics (reference src/lib.rs:478-490) */
    uint64_t non_zero, dimensions, iterations;
    int transposed;
    uint64_t lanczos_steps, ritz_values_stabilized, significant_values, singular_values;
    double end_interval[2];
    double kappa;
    uint32_t random_seed;
    /* trace, for step-by-step tests: T's diagonals, final ritz/bnd, counters */
    uint64_t trace_len; /* lanczos_steps */
    double *alf;        /* trace_len */
    double *bet;        /* trace_len + 1 */
    double *ritz;       /* trace_len */
    double *bnd;        /* trace_len */
    uint64_t n_opb;     /* svd_opb applications */
    uint64_t n_opa;     /* bare svd_opa applications (ritvec) */
    uint64_t n_purge_fired, n_purge_vecs, n_restarts;
    double t_total, t_opb, t_purge, t_imtql2, t_ritvec; /* seconds, CLOCK_MONOTONIC */
    char message[160];
} oracle_result_t;

/* For CSR/CSC: a = major offsets (len major+1), b = minor indices (nnz).
 * For COO: a = row indices (nnz), b = col indices (nnz).  All indices u64 (usize in the reference). */
int las2_oracle_svd(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                    const double *values, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                    double kappa, uint32_t random_seed, oracle_result_t **out);
void las2_oracle_free(oracle_result_t *r);

/* building blocks exposed for parity tests of individual kernels */
void las2_oracle_opa(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                     const double *values, const double *x, double *y, int transposed);
void las2_oracle_opb(int fmt, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a, const uint64_t *b,
                     const double *values, const double *x, double *y, double *temp, int transposed);
/* seconds per purge visit / per lanczos_step BLAS-1 chain / per contraction term on this host (bench.py's bounded sample) */
void las2_oracle_unit_costs(uint64_t n, uint64_t nvec, double out[3]);
int las2_oracle_imtqlb(uint64_t n, double *d, double *e, double *bnd);
int las2_oracle_imtql2(uint64_t nm, uint64_t n, double *d, double *e, double *z);
double las2_oracle_pythag(double a, double b);
void las2_oracle_rotate(double *a, uint64_t len, uint64_t x);
void oracle_start_vector(uint32_t seed, size_t n, double *out);
void oracle_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]);

#ifdef __cplusplus
}
#endif
#endif
