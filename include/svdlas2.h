/*
 * include/svdlas2.h -- C ABI of libsvdlas2_b200.so, the B200-native (sm_100a) LAS2 Lanczos SVD.
 *
 * This is the drop-in boundary for the hot path of dfarnham/svdlibrs (reference paths are into its
 * src/lib.rs):
 *
 *   reference interface                                   this ABI
 *   ----------------------------------------------------  --------------------------------------------
 *   svdLAS2(&dyn SMat, dimensions, iterations,             svdlas2_csr / svdlas2_csc / svdlas2_coo
 *           &end_interval, kappa, random_seed)  :570-577     (one per nalgebra-sparse input type, whose
 *   svd / svd_dim / svd_dim_seed  :499, 511, 524             raw arrays csr_data()/csc_data()/triplets
 *     (one-line wrappers with defaults :500, 512, 525)       are usize == uint64_t: zero-copy from Rust)
 *   SvdRec { d, ut, s, vt, diagnostics }  :455-461         svdlas2_result (library-owned host buffers)
 *   Diagnostics { ... 11 fields ... }     :478-490         svdlas2_diagnostics (same fields, same meaning)
 *   SvdLibError::{Imtqlb,Startv,Stpone,Imtql2,Las2}Error   return codes 1..5 + svdlas2_last_error()
 *     src/error.rs:8-26
 *   assert!/panic sites (:602-603, :1029, :1489, :1627)    SVDLAS2_ERR_REFERENCE_PANIC (never aborts)
 *   SMat::svd_opa / svd_opb  :443, :829-837                svdlas2_operator_opa / svdlas2_operator_opb
 *
 * Plain pointers and sizes only; nothing here depends on torch, the C++ runtime or the CUDA headers.
 * There is NO CPU fallback: every entry point that computes requires a CUDA device and fails with
 * SVDLAS2_ERR_CUDA when none is usable.
 */
#ifndef SVDLAS2_H
#define SVDLAS2_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVDLAS2_ABI_VERSION 2

#if defined(__GNUC__) || defined(__clang__)
#define SVDLAS2_API __attribute__((visibility("default")))
#else
#define SVDLAS2_API
#endif

/* ---- status codes ------------------------------------------------------------------------------ */
enum {
    SVDLAS2_OK = 0,
    SVDLAS2_ERR_IMTQLB = 1,          /* SvdLibError::ImtqlbError  (src/lib.rs:1016) */
    SVDLAS2_ERR_STARTV = 2,          /* SvdLibError::StartvError  (src/lib.rs:1128) */
    SVDLAS2_ERR_STPONE = 3,          /* SvdLibError::StponeError  (src/lib.rs:1183) */
    SVDLAS2_ERR_IMTQL2 = 4,          /* SvdLibError::Imtql2Error  (src/lib.rs:1612) */
    SVDLAS2_ERR_LAS2 = 5,            /* SvdLibError::Las2Error    (src/lib.rs:597)  */
    SVDLAS2_ERR_INVALID_ARGUMENT = 6,/* malformed input arrays / null pointers / index out of range */
    SVDLAS2_ERR_REFERENCE_PANIC = 7, /* the reference would panic here (e.g. error_bound step==0, :1489) */
    SVDLAS2_ERR_CUDA = 100,          /* CUDA runtime failure or no usable device (no CPU fallback) */
    SVDLAS2_ERR_COMM = 101,          /* NCCL / peer-memory failure */
    SVDLAS2_ERR_ALLOC = 102          /* host or device allocation failure */
};

/* ---- outputs ----------------------------------------------------------------------------------- */

/* Diagnostics, reference src/lib.rs:478-490 (field for field; note :648-649: significant_values is the
 * number of triplets RETURNED (== d) and singular_values the number ACCEPTED, as in the reference). */
typedef struct svdlas2_diagnostics {
    uint64_t non_zero;
    uint64_t dimensions;
    uint64_t iterations;
    int32_t transposed;
    uint64_t lanczos_steps;
    uint64_t ritz_values_stabilized;
    uint64_t significant_values;
    uint64_t singular_values;
    double end_interval[2];
    double kappa;
    uint32_t random_seed;
} svdlas2_diagnostics;

/* Side-channel observability (the reference has none; nothing is ever printed). Times in seconds. */
typedef struct svdlas2_profile {
    double t_total;        /* whole call, host clock */
    double t_upload;       /* H2D + device conversion to CSR(B), CSR(B') */
    double t_lanczos;      /* lanso: all Lanczos batches incl. host analysis */
    double t_ritvec;       /* imtql2 + contraction + sigma/U tail */
    double t_imtql2;       /* host tridiagonal eigen-decomposition inside ritvec */
    double t_download;     /* D2H of ut / vt */
    double t_opb_device;   /* summed CUDA-event time of the opb SpMV pairs (0 unless profiling is on) */
    double t_device;       /* CUDA-event time on the operator's stream from solve entry to the last D2H */
    double t_sync_wait;    /* host time blocked waiting for per-step scalars (inside t_lanczos) */
    uint64_t n_opb;        /* applications of B'(B x) */
    uint64_t n_spmv;       /* SpMV launches (2 per opb, +1 per returned vector) */
    uint64_t n_purge_fired;
    uint64_t n_purge_passes;
    uint64_t n_purge_vectors; /* stored vectors swept by purge (per pass) */
    uint64_t n_restarts;
    uint64_t n_kernel_launches;
    uint64_t n_host_syncs;
    uint64_t opb_algorithmic_bytes; /* 24*Z + 20*(M+N) + 8, SURVEY 8(d) */
    uint64_t h2d_bytes, d2h_bytes;
} svdlas2_profile;

/* SvdRec, reference src/lib.rs:455-461.  ut is d x ut_cols (ut_cols == rows(A)), vt is d x vt_cols
 * (vt_cols == cols(A)), both row-major exactly as Array::from_shape_vec((d, cols), value) at :638-640. */
typedef struct svdlas2_result {
    uint64_t d;
    uint64_t ut_cols;
    uint64_t vt_cols;
    double *ut;
    double *s;
    double *vt;
    svdlas2_diagnostics diagnostics;
    svdlas2_profile profile;
    /* trace of the Lanczos recurrence for step-by-step parity tests */
    uint64_t trace_len; /* == lanczos_steps */
    double *alf;        /* trace_len */
    double *bet;        /* trace_len + 1 */
    double *ritz;       /* trace_len */
    double *bnd;        /* trace_len */
    /* Sharded operators only: which of the d N-side vectors (rows of vt, or of ut on the transposed path) this rank's result
     * holds: rows [nside_row_begin, nside_row_begin + nside_row_count), stored from the array's first row on.  Everywhere
     * else: 0 and d.  (Knob "nside_mode": 0 = every rank downloads all of them, 1 = rank 0 only, 2 = the rows are dealt
     * out in contiguous blocks over the ranks, so that no rank moves more than 1/G of the replicated vectors over PCIe.) */
    uint64_t nside_row_begin;
    uint64_t nside_row_count;
} svdlas2_result;

/* ---- one-shot entry points (what a Rust shim's svdLAS2 binds) --------------- */

/* CsrMatrix<f64>::csr_data() -> (row_offsets[nrows+1], col_indices[nnz], values[nnz]); src/lib.rs:2083 */
SVDLAS2_API int svdlas2_csr(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *row_offsets,
                const uint64_t *col_indices, const double *values, uint64_t dimensions, uint64_t iterations,
                const double end_interval[2], double kappa, uint32_t random_seed, svdlas2_result **out);

/* CscMatrix<f64>::csc_data() -> (col_offsets[ncols+1], row_indices[nnz], values[nnz]); src/lib.rs:2047 */
SVDLAS2_API int svdlas2_csc(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *col_offsets,
                const uint64_t *row_indices, const double *values, uint64_t dimensions, uint64_t iterations,
                const double end_interval[2], double kappa, uint32_t random_seed, svdlas2_result **out);

/* CooMatrix<f64> triplets (unsorted, duplicates allowed and effectively summed); src/lib.rs:2121-2127 */
SVDLAS2_API int svdlas2_coo(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *row_indices,
                const uint64_t *col_indices, const double *values, uint64_t dimensions, uint64_t iterations,
                const double end_interval[2], double kappa, uint32_t random_seed, svdlas2_result **out);

SVDLAS2_API void svdlas2_free(svdlas2_result *result);

/* Message of the last failure on the calling thread ("" if none).  Valid until the next call on it. */
SVDLAS2_API const char *svdlas2_last_error(void);

SVDLAS2_API int svdlas2_abi_version(void);

/* The library keeps the result buffers it hands out (registered host memory) and its device scratch (the
 * stream-ordered memory pool of the current device) cached between calls, because creating either costs more than
 * a small solve.  This gives everything that is not in use back to the operating system / the driver -- e.g. before
 * another library in the same process needs the GPU's memory.  Returns 0. */
SVDLAS2_API int svdlas2_trim_caches(void);

/* ---- resident-operator entry points ------------------------------------------------------------- *
 * The same path split in two so a caller (an `impl SMat`, a benchmark, a parity test) can keep the
 * operator in HBM: create = the reference's implicit "borrow the matrix", solve = lanso + ritvec.       */

typedef struct svdlas2_operator svdlas2_operator;

enum { SVDLAS2_FMT_CSR = 0, SVDLAS2_FMT_CSC = 1, SVDLAS2_FMT_COO = 2 };

/* a/b as in the one-shot calls (offsets+indices, or row+col triplets).  device < 0 -> current device.
 * Orientation (src/lib.rs:606) is decided here: B = A, or A' when ncols >= 1.2*nrows. */
SVDLAS2_API int svdlas2_operator_create(int format, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a,
                            const uint64_t *b, const double *values, int device, svdlas2_operator **op);
SVDLAS2_API void svdlas2_operator_destroy(svdlas2_operator *op);

/* svdLAS2 (src/lib.rs:570-655) on a resident operator.  random_seed == 0 draws a seed from OS entropy like the
 * reference (:578-581) and reports it in Diagnostics; on a sharded operator rank 0 draws it and shares it with the
 * other ranks through the communicator, so all ranks build the same start vector. */
SVDLAS2_API int svdlas2_operator_solve(svdlas2_operator *op, uint64_t dimensions, uint64_t iterations,
                           const double end_interval[2], double kappa, uint32_t random_seed,
                           svdlas2_result **out);

/* SMat::svd_opa (src/lib.rs:443): y = A x (transposed == 0, x: cols(A), y: rows(A)) or y = A' x. Host buffers. */
SVDLAS2_API int svdlas2_operator_opa(svdlas2_operator *op, const double *x, double *y, int transposed);
/* svd_opb (src/lib.rs:829-837) on the oriented operator: y = B'(B x); x, y: N doubles. Host buffers. */
SVDLAS2_API int svdlas2_operator_opb(svdlas2_operator *op, const double *x, double *y);

/* shape of the oriented operator B (M x N) and nnz; transposed flag as in Diagnostics */
SVDLAS2_API int svdlas2_operator_shape(const svdlas2_operator *op, uint64_t *m, uint64_t *n, uint64_t *nnz, int *transposed);

/* Device-timed opb for the roofline: `reps` back-to-back applications on the operator's stream between two
 * CUDA events, x resident in HBM.  flush_l2 != 0 writes a >L2-sized buffer before each application (outside
 * the per-application event pair).  Returns mean milliseconds per opb and per SpMV half. */
SVDLAS2_API int svdlas2_operator_time_opb(svdlas2_operator *op, int reps, int flush_l2, double *ms_per_opb,
                              double *ms_spmv_b, double *ms_spmv_bt);

/* knobs: "profile" (0/1: per-opb CUDA events); "nside_mode" (sharded operators: the N-side singular vectors are replicated
 * over the ranks; 0 = every rank downloads all of them, 1 = only rank 0 does and the other ranks return an N-side array of
 * zero columns, 2 = every rank downloads its contiguous block of them, see svdlas2_result.nside_row_begin; the M-side
 * vectors always come out row-sliced per rank); "nside_on_rank0_only" (0/1) = nside_mode 0 / 1; tuning / test overrides
 * of the SpMV: "spmv_variant" (-1 auto, 5 = warp-chunk stream, 0..3 = the tile kernels on the natural CSR arrays),
 * "spmv_hot", "spmv_shape", "spmv_pdl", "spmv_carveout", "spmv_smem_pad_kb".  Every knob is stored in the operator it
 * is set on; nothing is process-wide.
 * Threading: the library is re-entrant across operators (no global mutable state besides lazily created caches that
 * are guarded); ONE operator, and ONE communicator, serve one thread at a time (they own a stream, scratch buffers
 * and collective staging), exactly like a `&mut` borrow in the reference's language. */
SVDLAS2_API int svdlas2_operator_set(svdlas2_operator *op, const char *knob, int64_t value);

/* read-only facts about a resident operator (reporting): "relabelled" (1 when the N side was re-numbered by popularity at
 * upload), "ring" (1 when the sharded exchange runs as the peer-memory kernel, 0 = ncclAllReduce / single GPU),
 * "ring_calls", "ring_two_shot_calls", "ring_bytes_per_call" (bytes this rank moves over NVLink per exchange),
 * "hot_fraction" (share of B's gathers served from the shared-memory prefix).  Unknown keys return an error. */
SVDLAS2_API int svdlas2_operator_info(const svdlas2_operator *op, const char *key, double *value);

/* ---- multi-GPU, one process per GPU (rows of the oriented operator B sharded by nnz) -- */

#define SVDLAS2_COMM_ID_BYTES 128
/* rank 0 fills an opaque id, the caller broadcasts it (torch.distributed / MPI / a file), then every rank
 * creates its communicator with the same id.  A communicator is long-lived and serves any number of shards. */
SVDLAS2_API int svdlas2_comm_unique_id(uint8_t id[SVDLAS2_COMM_ID_BYTES]);

typedef struct svdlas2_comm svdlas2_comm;
SVDLAS2_API int svdlas2_comm_create(int rank, int world_size, const uint8_t id[SVDLAS2_COMM_ID_BYTES], int device,
                        svdlas2_comm **comm);
SVDLAS2_API void svdlas2_comm_destroy(svdlas2_comm *comm);

/* Row shard [row_begin, row_end) of B in CSR form: local_offsets has (row_end-row_begin)+1 entries starting
 * at 0, local_indices are column indices < n.  m_global/n are B's full shape.  `b_is_a_transposed` records
 * the orientation the caller applied (Diagnostics.transposed; ut/vt are swapped back accordingly).
 * `comm` (borrowed, must outlive the operator) may be NULL for a single-rank run.
 * After solve each rank's result holds all of the N-side vectors and ITS OWN row slice of the M-side ones
 * (ut_cols or vt_cols == row_end-row_begin), so no collective is needed for them. */
SVDLAS2_API int svdlas2_operator_create_shard(uint64_t m_global, uint64_t n, uint64_t row_begin, uint64_t row_end,
                                  uint64_t local_nnz, const uint64_t *local_offsets,
                                  const uint64_t *local_indices, const double *values, int b_is_a_transposed,
                                  uint64_t global_nnz, svdlas2_comm *comm, int device, svdlas2_operator **op);

#ifdef __cplusplus
}
#endif
#endif /* SVDLAS2_H */
