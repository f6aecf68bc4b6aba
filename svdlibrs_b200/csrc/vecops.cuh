// vecops.cuh -- fused BLAS-1 on the device-resident Lanczos vectors, and the blocked panel operations.
//
// Replaces the reference's svd_daxpy / svd_ddot / svd_norm / svd_datx / svd_dscal loops
// (src/lib.rs:840-844, 893-914) as they are chained in stpone (:1187-1198), lanczos_step (:1274-1304),
// purge (:1372-1388) and startv (:1132-1144).
//
// Conventions
//  * every N-vector is allocated with `ld` = N rounded up to 32 doubles and its padding is kept at 0.0,
//    so kernels run over ld/2 double2 elements without tail handling;
//  * reductions are two-stage and deterministic: each CTA writes one partial (fixed-order shuffle tree),
//    and whoever needs the scalar sums the partials again with one fixed-order warp reduction
//    (cta_sum_partials) -- producers and consumers therefore agree bit for bit, on every GPU of a
//    replicated multi-GPU run, with no atomics;
//  * element-wise updates use separate multiply and add (no FMA contraction), as the Rust reference does.
#pragma once

#include "util.cuh"

namespace las2 {

constexpr int kVecThreads = 256;
constexpr int kMaxParts = 592; // 148 SMs x 4: upper bound on the grid of any reducing kernel

// A scalar that is either known on the host (imm) or still sitting in HBM as per-CTA partial sums.
struct Coef {
    const double *parts;
    int nparts;
    double imm;
    static Coef value(double v) { return Coef{nullptr, 0, v}; }
    static Coef partials(const double *p, int n) { return Coef{p, n, 0.0}; }
};

struct PartialSlot {
    double *parts; // kMaxParts doubles
    int nparts;    // set by the producing launch
};

unsigned vec_grid(u64 ld);

// dst = alpha * src                                   (svd_datx / svd_dscal, :903-914)
void vec_scale(double *dst, const double *src, double alpha, u64 ld, cudaStream_t s);
// out = a . b                                         (svd_ddot, :893-895)
void vec_dot(const double *a, const double *b, u64 ld, PartialSlot &out, cudaStream_t s);
// r -= coef * x ;  out = r . y   (y == nullptr: out = r . r)       (svd_daxpy + svd_ddot pairs)
void vec_axpy_dot(double *r, const double *x, Coef coef, const double *y, u64 ld, PartialSlot &out, cudaStream_t s);
// r -= coef * x                                       (svd_daxpy, :840-844)
void vec_axpy(double *r, const double *x, Coef coef, u64 ld, cudaStream_t s);

// The local re-orthogonalisation chain of one Lanczos step (reference src/lib.rs:1279-1304) as ONE launch:
//   phase k:  c_k = canonical sum of the previous phase's partials (phase 0: the alf partials of the opb epilogue) ;
//             r -= c_k * x_k ;  partials of  r . y_k   (y_k == nullptr: r . r)
// with a grid-wide barrier between the phases, instead of 3-5 dependent launches of k_axpy_dot plus k_gather_scalars.  The
// last CTA out writes the scalars the host needs (alf, t1, t2, |r|^2 -- out[0..4) in that order, see `pick`) to pinned host
// memory, followed by a sequence number the host polls: no cudaStreamSynchronize per step.  Optionally also writes the next
// Lanczos vector q_next = r / sqrt(|r|^2) (the host discards it when purge / a restart changes r).
struct StepChain {
    int nphases;              // 3 .. 6
    const double *x[6];       // axpy vectors
    const double *y[6];       // dot vectors (nullptr: the norm)
    int pick[4];              // which phase's coefficient goes to out[0], out[1], out[2]; out[3] is always the final norm^2
    const double *alf_parts;  // partials of the coefficient of phase 0 (nullptr: the coefficient is coef0, known on the host --
    int alf_nparts;           // the opb epilogue r -= rnm q_{j-1} folded into the chain as its first phase)
    double coef0;
    double *q_next;           // may be nullptr
};
// The whole Lanczos step of a SMALL operator in the same launch (single GPU, both matrices in plain CSR, short rows): two more
// leading phases compute t = B x and r = B' t - coef0 * qprev with the partials of r . x (svd_opb, src/lib.rs:829-837, plus
// :1276-1278), so that a step is ONE launch and one polled flag instead of 5-6 launches.  Rows are handed to sub-warps of
// `lanes` threads (2..32, chosen from the mean row length); a lane adds its strided elements in order (fma), a fixed xor
// tree joins the lanes: deterministic, independent of the grid.  The chain then starts at phase 1 (phase 0 of `StepChain` must
// be the folded epilogue: x[0] = qprev, y[0] = x, alf_parts = nullptr).
struct StepSpmv {
    const u32 *b_off, *b_idx;   const double *b_val;   u32 b_rows;  int b_lanes;   // B  (M x N)
    const u32 *bt_off, *bt_idx; const double *bt_val;  u32 bt_rows; int bt_lanes;  // B' (N x M)
    double *t;                  // M doubles of scratch
    int solo_tail;              // set by vec_step_fused: the vector phases run on one CTA (short vectors)
};
// sub-warp width for rows of the given mean length / the longest row such a launch still takes
int step_spmv_lanes(double mean_row_len);
inline u32 step_spmv_max_row(int lanes) { return 64u * (u32)lanes; }
// longest row of a CSR matrix (one small kernel + one synchronising read-back)
u32 csr_max_row_len(const u32 *off, u32 rows, cudaStream_t s);
void vec_step_fused(double *r, u64 ld, const StepChain &c, const StepSpmv &m, double *work, double *host_out, unsigned long long seq,
                    cudaStream_t s);
// work: 6 * kMaxParts doubles of partial scratch + 2 unsigned ints (ZEROED once by the caller: the launch's grid barrier);
// host_out: 4 doubles + 1 sequence word (pinned).  The grid is at most one CTA per SM.
void vec_step_chain(double *r, u64 ld, const StepChain &c, double *work, double *host_out, unsigned long long seq, cudaStream_t s);
unsigned step_chain_max_grid(); // co-resident CTAs of the chain kernel on the current device (0: cooperative launch unavailable)

// Sum up to 8 partial slots with the canonical reduction and write the scalars to `out` (pinned host memory
// or device memory).
struct GatherList {
    const double *parts[8];
    int nparts[8];
    int count;
};
void gather_scalars(const GatherList &g, double *out, cudaStream_t s);

// ---- panel operations on Q (column-major, leading dimension ld), columns [col0, col0+ncols) ----------
// Classical Gram-Schmidt sweep of the pair (a, b) against the stored Lanczos vectors, as purge performs it
// with its frozen snapshots w3/w4 (src/lib.rs:1374-1382):
//      ca = Q'a, cb = Q'b ;  a -= Q ca, b -= Q cb ;  sum|ca|, sum|cb| ;  dot_out = a . b (updated)
// b may be nullptr (single right-hand side: the restart sweep of startv, :1132-1135).
struct PanelWork {
    DevBuf<double> partials; // row-block partial dots
    DevBuf<double> coeffs;   // 2 * ncols
    DevBuf<double> dotparts; // one per update CTA when that exceeds kMaxParts
    size_t cap_cols = 0, cap_blocks = 0;
};
// abs_a / abs_b receive the partials of sum|ca| / sum|cb| ; dot_out (optional, needs b) those of a . b.
// With a communicator (replicated panel, sharded operator) and a sweep big enough to pay for two small collectives,
// the ROWS of the sweep are split over the ranks: every rank forms the coefficient partials of
// its row blocks, one all-reduce of 2*ncols doubles completes them, every rank updates its rows of (a, b) and one
// all-gather hands everybody the whole pair again -- bit-identical on all ranks, 1/G of the panel traffic each.
class Comm;
void panel_cgs2(const double *Q, u64 ld, u64 n, u32 col0, u32 ncols, double *a, double *b, PanelWork &work,
                PartialSlot *abs_a, PartialSlot *abs_b, PartialSlot *dot_out, cudaStream_t s, u64 *launches,
                cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_mid = nullptr, cudaEvent_t ev_end = nullptr,
                Comm *comm = nullptr, u64 *sharded_sweeps = nullptr);

// Ritz back-transform (ritvec, src/lib.rs:1807-1817): V[:, c] = sum_k Q[:, k] * C[k, c], k < js, c < d.
// C is js x d row-major on the device; V is column-major with leading dimension ld (each Ritz vector contiguous).
void ritz_contract(const double *Q, u64 ld, u64 n, u32 js, const double *C, u32 d, double *V, cudaStream_t s,
                   u64 *launches, Comm *comm = nullptr, bool *sharded = nullptr);
extern int g_ritz_variant; // test / tuning hook: -1 auto (SVDLAS2_RITZ), 0 FMA pipe, 1 DMMA

} // namespace las2
