// las2.hpp -- the resident operator and the LAS2 driver (host control + kernel launches).
#pragma once

#include <memory>
#include <vector>

#include "comm.hpp"
#include "peer_ring.cuh"
#include "device_csr.cuh"
#include "host_math.hpp"
#include "vecops.cuh"

namespace las2 {

// The operator B (M x N, M >= N "mostly": B = A, or A' when cols(A) >= 1.2 rows(A), src/lib.rs:606-608),
// resident in HBM as CSR(B) + CSR(B').  In a multi-GPU run each rank holds a contiguous row range of B.
struct Operator {
    int device = 0;
    cudaStream_t stream = nullptr;
    // shape of the caller's matrix A and of the oriented operator
    u64 a_rows = 0, a_cols = 0, nnz_global = 0;
    bool transposed = false;
    u64 M_global = 0; // rows of B over all ranks
    u64 M = 0;        // rows of B held here
    u64 N = 0;
    u64 row_begin = 0;
    DeviceCsr B, BT; // BT is N x M (local)
    Relabel relabel;  // N-side ids re-assigned at upload (device_csr.cuh); vectors inside the solver are in the NEW order
    Comm *comm = nullptr; // borrowed; null on a single GPU
    // the exchange step of a sharded opb over NVLink peer memory, fused with the first update of the Lanczos step
    // (peer_ring.cuh); null on a single GPU, when peer memory cannot be mapped, or with SVDLAS2_RING=0 (then: ncclAllReduce).
    // Borrowed from the communicator, which keeps ONE ring for all the operators it serves (mapping peer memory costs far
    // more than an upload of a small shard; an operator created per call must not pay it every time).
    PeerRing *ring = nullptr;
    // scratch for opa/opb/time_opb
    DevBuf<double> tmp_m, tmp_x, tmp_y;
    DevBuf<char> l2_flush;
    // knobs
    bool profile = false;
    int nside_mode = 0; // sharded runs, download of the replicated N-side vectors: 0 every rank, 1 rank 0 only, 2 dealt out by rows
    // rows [b, e) of the d N-side vectors this rank brings to the host
    void nside_rows(u64 d, u64 *b, u64 *e) const {
        *b = 0; *e = d;
        if (!comm) return;
        if (nside_mode == 1 && comm->rank() != 0) *e = 0;
        if (nside_mode == 2) { *b = d * (u64)comm->rank() / (u64)comm->world(); *e = d * (u64)(comm->rank() + 1) / (u64)comm->world(); }
    }
    // bookkeeping
    UploadStats upload;
    double t_upload = 0.0;

    ~Operator();
    u64 opb_bytes() const { return B.spmv_bytes() + BT.spmv_bytes(); } // 24 Z + 20 (M + N) + 8
    // y = B'(B x) on the stream; x, y device N-vectors (y != x); t scratch of M doubles
    void opb(const double *x, double *y, double *t, u64 *launches);
    // r = B'(B x) - rnm * qprev with the partials of r . x in *alf (reference :1276-1278 in one go); qprev == nullptr: r =
    // B'(B x) only.  With a ring, r MUST be ring->r_mine() (the peers push their slices into it).
    void opb_step(const double *x, const double *qprev, double rnm, double *r, double *t, PartialSlot *alf, u64 ld, u64 *launches);
};

// drops the process-wide spare storage of imtql2's rotation plan (svdlas2_trim_caches)
void plan_cache_trim();

struct SolveOutput; // svdlas2_result with owned buffers (abi.cu)

// svdLAS2 (src/lib.rs:570-655) on a resident operator; fills *res (already calloc'ed). Throws las2::Error.
void solve(Operator &op, u64 dimensions, u64 iterations, const double end_interval[2], double kappa,
           uint32_t random_seed, svdlas2_result *res);

} // namespace las2
