/*
 * include/svdlas2_testing.h -- test hooks into the HOST-side logic of libsvdlas2_b200.so (no GPU needed).
 * They exist so that the CPU test-suite can compare the product's control-flow arithmetic with the oracle bit
 * for bit; they are not part of the drop-in boundary (include/svdlas2.h).
 */
#ifndef SVDLAS2_TESTING_H
#define SVDLAS2_TESTING_H
#include "svdlas2.h"

#ifdef __cplusplus
extern "C" {
#endif

/* first n samples of StdRng(seed).gen_range(-1.0..1.0)  (reference src/lib.rs:1109-1114) */
SVDLAS2_API void svdlas2_test_start_vector(uint32_t seed, uint64_t n, double *out);
/* svd_pythag (src/lib.rs:873-890) */
SVDLAS2_API double svdlas2_test_hypot(double a, double b);
/* imtqlb (src/lib.rs:962-1068): d, e, bnd overwritten; returns a status code */
SVDLAS2_API int svdlas2_test_ql_values(uint64_t n, double *d, double *e, double *bnd);
/* imtql2 (src/lib.rs:1576-1693) through the rotation plan: d, e overwritten, z (n x n row-major) receives the
 * eigenvectors exactly as the reference's in-place update of an identity would (rotations replayed on the CPU
 * here; the product replays the same plan on the GPU).  n_rotations / n_sweeps are optional outputs. */
SVDLAS2_API int svdlas2_test_ql_vectors(uint64_t n, double *d, double *e, double *z, uint64_t *n_rotations,
                                        uint64_t *n_sweeps);
/* The rotation plan recorded inside imtqlb (what the last analysis of T hands to ritvec) against the plan of imtql2's own
 * recurrence on the same (d, e): returns 1 when rotations, sweeps and column permutation are bit-identical (or both fail to
 * converge) and recording leaves imtqlb's outputs untouched, 0 when the plans differ, negative on inconsistent status codes. */
SVDLAS2_API int svdlas2_test_ql_plans_agree(uint64_t n, const double *d, const double *e, uint64_t *n_rotations);
/* error_bound (src/lib.rs:1480-1532): returns status; neig and the updated enough flag through pointers */
SVDLAS2_API int svdlas2_test_refine_bounds(int *enough, double endl, double endr, double *ritz, double *bnd,
                                           uint64_t step, double tol, uint64_t *neig);
/* insert_sort (src/lib.rs:815-825) */
SVDLAS2_API void svdlas2_test_cosort(uint64_t n, double *keys, double *vals);
/* Host result pool (host_pool.cpp): acquires a block of `bytes`, writes its first and last byte, releases it and trims the
 * cache.  *node receives the memory node the block was steered to (the node of the current CUDA device; -1 when there is no
 * device or the topology is not exposed).  Works without a GPU (the block is then pageable).  Returns a status code. */
SVDLAS2_API int svdlas2_test_host_block(uint64_t bytes, int *node);
/* GPU tuning hook (needs a device): device-times `reps` reorthogonalisation passes of purge (src/lib.rs:1372-1388) of a
 * pair (q, r) against `ncols` synthetic Lanczos vectors of length n; outputs milliseconds per pass spent in the
 * coefficient kernels (c = Q'[q r]) and in the update kernel ([q r] -= Q c).  Returns a status code. */
SVDLAS2_API int svdlas2_test_time_panel(uint64_t n, uint32_t ncols, int reps, int device, double *ms_dots,
                                        double *ms_update);

/* GPU test hook for the 4-right-hand-side product of the ritvec tail (k_spmm4_chunk): Y[r] = B X[r] for r < nv <= 4 on the
 * oriented operator of `op`; X holds nv vectors of N doubles, Y receives nv vectors of M doubles, sumsq[r] = |Y[r]|^2 as the
 * tail computes it.  Returns SVDLAS2_ERR_INVALID_ARGUMENT when the operator's B is not on the chunk stream. */
SVDLAS2_API int svdlas2_test_spmm4(svdlas2_operator *op, uint32_t nv, const double *X, double *Y, double *sumsq);

/* GPU tuning / test hook for the Ritz back-transform V = Q C of ritvec (src/lib.rs:1807-1817): Q is n x js (pseudo-random), C is
 * js x d; both kernels run `reps` times: ms[0] = FMA-pipe register tile (k_ritz_contract), ms[1] = DMMA tile kernel
 * (k_ritz_contract_dmma); max_diff = largest |V_fma - V_dmma| relative to the largest |V|; max_err = largest deviation of the
 * DMMA result from a host-computed double-double reference on a sample of entries, same scale. */
SVDLAS2_API int svdlas2_test_time_ritz(uint64_t n, uint32_t js, uint32_t d, int reps, int device, double ms[2], double *max_diff,
                                       double *max_err);
/* GPU test / tuning hook for the device half of imtql2 (src/lib.rs:1646-1652): runs the host recurrences on (d, e) (both
 * overwritten, d = ascending eigenvalues), replays the recorded rotations on the GPU `reps` times and returns z (n x n row-major,
 * logical column order) exactly as svdlas2_test_ql_vectors does on the CPU -- the two must agree BIT FOR BIT.  variant: 0 = one
 * replaying warp per CTA, 1 = the sweeps pipelined over the warps of the CTA, -1 = the library's default.  ms = wall time of one
 * replay incl. the upload of the plan (reps >= 2). */
SVDLAS2_API int svdlas2_test_ql_replay(uint64_t n, double *d, double *e, double *z, int variant, int reps, int device, double *ms,
                                       uint64_t *n_rotations);
/* GPU test hook for the device allocator behind every buffer of the library (util.cu): large blocks are parked on release and
 * re-issued by exact fit; an allocation that does not fit next to what is parked gives the parked blocks back and succeeds.
 * out[0] = bytes parked after releasing a block of 30 % of the free memory, out[1] = 1 when a slightly smaller request got that very
 * block back, out[2] = 1 when a request of 80 % of the free memory then succeeded, out[3] = bytes parked after the final trim (0). */
SVDLAS2_API int svdlas2_test_device_parking(int device, double out[4]);
/* How the upload laid the operator out: out[0] = 1 when the N side was relabelled by popularity, out[1] = share of B's column
 * indices inside the shared-memory prefix (hot_fraction), out[2] = size of that prefix, out[3] / out[4] = mode of B's / B''s
 * first stream and out[5] of B''s second one (-1 none, 0 plain, 1 plain + hot prefix, 2 slab-local, 3 wide slabs),
 * out[6] = first row of B''s second stream, out[7] = number of slabs of B''s first stream. */
SVDLAS2_API int svdlas2_test_operator_layout(const svdlas2_operator *op, double out[8]);
/* GPU test hook for the peer-memory exchange kernel of the sharded svd_opb (k_allreduce_epilogue, peer_ring.cu): G ranks
 * emulated by ONE launch on one GPU (blockIdx.y = rank, G * ctas_per_rank <= 148 so that all CTAs are co-resident).  ys: G
 * partial vectors of n doubles; out: what every rank ends up with, G vectors of n doubles (must all be equal:
 * sum_g ys[g] - rnm * qprev, partials added in rank order); dots[g] = rank g's value of out . q.  n % 32 == 0. */
SVDLAS2_API int svdlas2_test_ring_emulation(int G, uint64_t n, const double *ys, const double *qprev, double rnm,
                                            const double *q, int two_shot, int ctas_per_rank, double *out, double *dots);

#ifdef __cplusplus
}
#endif
#endif
