// peer_ring.cuh -- the exchange step of the sharded svd_opb as ONE kernel over NVLink peer memory, fused with the first
// vector update of the Lanczos step that follows it (SURVEY 2.3 "allreduce_epilogue", 5; replaces ncclAllReduce + k_axpy_dot):
//
//      r = sum_g y_g  -  rnm * q_{j-1} ,   alf partials of  r . q_j         (reference src/lib.rs:835-836 + :1277-1278)
//
// Every rank's partial product y_g = B_g'(B_g x) lives in a buffer the other ranks have mapped (CUDA IPC, one process per
// GPU).  The partials are ALWAYS added in rank order 0, 1, ..., G-1, so the result is bit-identical on every rank and from
// run to run, whatever the timing -- no NCCL algorithm / protocol choice is involved.
//   * one-shot (N * 8 B <= 2 MB): every rank reads all G partials of every element over NVLink and forms the whole of r and
//     of the dot product itself; one cross-GPU flag barrier per call, y double-buffered by call parity.
//   * two-shot (larger N): rank g reduces only its slice [N g / G, N (g+1) / G) (reads (G-1)/G * N / G ... of remote data),
//     applies the epilogue, PUSHES the finished slice of r and its dot partials into every peer's buffers, and a second flag
//     barrier closes the call.  Per rank and call: 2 (G-1) N / G doubles cross NVLink instead of 2 (G-1) N / G in a ring
//     all-reduce's G-1 latency-bound steps.
#pragma once

#include "util.cuh"
#include "vecops.cuh"

namespace las2 {

class Comm;

constexpr int kRingMaxRanks = 8;
constexpr int kRingCtas = kNumSMs; // one CTA per SM; also the number of dot partials a rank contributes
constexpr u64 kRingSmallCap = 16384; // doubles: the small channel (reorthogonalisation coefficients, squared norms, counts)

struct PeerRing {
    int G = 1, g = 0;
    u64 cap = 0;           // doubles per vector buffer
    void *base = nullptr;  // this rank's symmetric block (cudaMalloc: IPC-capable)
    void *peer_base[kRingMaxRanks] = {};
    double *y[2][kRingMaxRanks] = {};             // partial products, by call parity
    double *r[kRingMaxRanks] = {};                // the reduced vector (w0 of the reference) on every rank
    double *parts[kRingMaxRanks] = {};            // G * kRingCtas dot partials on every rank
    double *ys[2][kRingMaxRanks] = {};            // small channel: inputs by call parity ...
    double *rs[kRingMaxRanks] = {};               // ... and the result
    unsigned long long *flags[kRingMaxRanks] = {}; // 4 x kRingMaxRanks flag words on every rank
    unsigned long long seq = 0;                   // call number (same on every rank: the ranks run in lockstep)
    u64 two_shot_min_bytes = 2u << 20;
    // statistics
    u64 calls = 0, two_shot_calls = 0;

    ~PeerRing();
    // un-map the peers' blocks (idempotent).  Owners do: close_peers() on every rank, a barrier, THEN destroy the ring --
    // an exporter must not free its block while another process still has it open
    void close_peers();
    // zero this rank's vector buffers (a new operator attaches: the padding beyond its N must read 0.0)
    void reset_vectors(cudaStream_t s);
    // collective: every rank of `comm` calls it with the same n.  Throws Error(SVDLAS2_ERR_COMM) when peer memory is not
    // available between the ranks (the caller then stays on the NCCL path).
    static PeerRing *create(Comm *comm, u64 n_doubles);
    double *y_mine() const { return y[(seq + 1) & 1][g]; } // buffer the NEXT call reduces: the SpMV writes here
    double *r_mine() const { return r[g]; }
    // r = sum_g y_g - rnm * qprev (qprev == nullptr: plain sum), dot partials of r . q (q == nullptr: none) into `dot`
    // (dot->parts then points into the ring's partial buffer).  ld2 = length in double2 units.
    void reduce(u64 ld, const double *qprev, double rnm, const double *q, PartialSlot *dot, cudaStream_t s);
    // in-place sum over the ranks of a SHORT device vector (n <= kRingSmallCap), added in rank order like everything else:
    // every floating-point reduction across ranks is then independent of NCCL's algorithm / protocol choices
    void allreduce_small(double *buf, u64 n, cudaStream_t s);
};

void ring_quiesce(PeerRing &R, cudaStream_t s);
// single-GPU emulation of G ranks in ONE launch (tests of the protocol and of the arithmetic; see peer_ring.cu)
void ring_emulate(int G, u64 ld, const double *ys, const double *qprev, double rnm, const double *q, int two_shot, int ctas_per_rank,
                  double *out, double *dots_parts, int *nparts, cudaStream_t s);

} // namespace las2
