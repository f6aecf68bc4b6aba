// peer_ring.cu -- see peer_ring.cuh.  One kernel per svd_opb on every GPU: cross-GPU flag barrier, rank-ordered sum of the G
// partial products over NVLink peer loads, the epilogue r -= rnm * q_{j-1} and the partials of alf = r . q_j, and (two-shot)
// the push of the finished slice into every peer.
#include "peer_ring.cuh"

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "comm.hpp"

namespace las2 {

namespace {

constexpr int kRingThreads = 512;
constexpr int kFlagWords = 4 * kRingMaxRanks; // [0,8): call barrier, [8,16): end-of-call barrier, [16,24): quiesce, [24]: CTA counter

struct RingArgs {
    const double2 *y[kRingMaxRanks];
    double2 *r[kRingMaxRanks];
    double *parts[kRingMaxRanks];
    unsigned long long *flags[kRingMaxRanks];
    int G, g;
    unsigned long long seq;        // barrier generation (monotonic over all calls of the ring)
    unsigned long long ctr_target; // value of this rank's CTA counter once every CTA of THIS call has finished (two-shot)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// peer data: read once, never cached in this GPU's L1
__device__ __forceinline__ double2 ld_peer(const double2 *p) {
    double2 v;
    asm volatile("ld.global.relaxed.sys.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

// wait until every rank has signalled `seq` in word block `blk` of MY flag array; CTA `signaller` of the grid signals the peers
__device__ __forceinline__ void ring_barrier(const RingArgs &A, int blk, bool signal) {
    const int h = threadIdx.x;
    if (h < A.G && h != A.g) {
        if (signal) {
            __threadfence_system();
            st_release_sys(&A.flags[h][blk * kRingMaxRanks + A.g], A.seq);
        }
        const unsigned long long *mine = &A.flags[A.g][blk * kRingMaxRanks + h];
        while (ld_acquire_sys(mine) < A.seq) { }
    }
    __syncthreads();
}

// A real run passes its RingArgs by value (gridDim.y == 1); the single-GPU emulation passes one RingArgs per emulated rank
// (blockIdx.y) through memory.
template <bool TWO_SHOT, bool EMULATED>
__global__ void __launch_bounds__(kRingThreads)
k_allreduce_epilogue(const __grid_constant__ RingArgs mine, const RingArgs *__restrict__ args, u64 n2,
                     const double2 *__restrict__ qprev, double rnm, const double2 *__restrict__ q, int want_dot) {
    __shared__ RingArgs A;
    __shared__ double sm[kRingThreads / 32];
    __shared__ int last_cta;
    if (threadIdx.x == 0) A = EMULATED ? args[blockIdx.y] : mine;
    __syncthreads();
    const int G = A.G, g = A.g;
    ring_barrier(A, 0, blockIdx.x == 0); // every rank's partial product is complete

    u64 lo = 0, hi = n2;
    if (TWO_SHOT) { lo = n2 * (u64)g / (u64)G; hi = n2 * (u64)(g + 1) / (u64)G; }
    const double nr = -rnm;
    double acc = 0.0;
    for (u64 i = lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (u64)gridDim.x * blockDim.x) {
        double2 part[kRingMaxRanks];
#pragma unroll
        for (int h = 0; h < kRingMaxRanks; h++)
            if (h < G) part[h] = (h == g) ? A.y[h][i] : ld_peer(A.y[h] + i);
        double2 s = part[0];
#pragma unroll
        for (int h = 1; h < kRingMaxRanks; h++)
            if (h < G) { s.x += part[h].x; s.y += part[h].y; } // rank order, whatever the timing: identical on every rank
        if (qprev) {
            const double2 p = qprev[i];
            s.x = __dadd_rn(s.x, __dmul_rn(nr, p.x)); // svd_daxpy(-rnm, w2, w0), reference :1277
            s.y = __dadd_rn(s.y, __dmul_rn(nr, p.y));
        }
        if (TWO_SHOT) {
#pragma unroll
            for (int h = 0; h < kRingMaxRanks; h++)
                if (h < G) A.r[h][i] = s; // my own copy and the push into every peer
        } else {
            A.r[g][i] = s;
        }
        if (want_dot) {
            const double2 qv = q[i];
            acc = fma(s.x, qv.x, acc); // alf = w0 . w3, :1278
            acc = fma(s.y, qv.y, acc);
        }
    }
    if (want_dot) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        acc = warp_sum(acc);
        if (lane == 0) sm[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kRingThreads / 32; w++) t += sm[w];
            if (TWO_SHOT) {
                for (int h = 0; h < G; h++) A.parts[h][(u64)g * gridDim.x + blockIdx.x] = t;
            } else {
                A.parts[g][blockIdx.x] = t;
            }
        }
    }
    if (TWO_SHOT) {
        // the last CTA of this rank to finish tells the peers that everything this rank pushes has landed, and leaves only
        // once every peer has said the same: the kernel's completion then means "r and the partials are whole here"
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long *ctr = &A.flags[g][3 * kRingMaxRanks];
            const unsigned long long done = atomicAdd(ctr, 1ull) + 1ull;
            last_cta = (done == A.ctr_target) ? 1 : 0;
            __threadfence();
        }
        __syncthreads();
        if (last_cta) ring_barrier(A, 1, true);
    }
}

// all ranks have finished every access to peer memory (end of a solve: buffers may be freed / reused afterwards)
template <bool EMULATED>
__global__ void k_ring_quiesce(const __grid_constant__ RingArgs mine, const RingArgs *__restrict__ args) {
    __shared__ RingArgs A;
    if (threadIdx.x == 0) A = EMULATED ? args[blockIdx.y] : mine;
    __syncthreads();
    ring_barrier(A, 2, true);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace

static void fill_args(const PeerRing &R, RingArgs &a, unsigned long long seq, bool small = false) {
    std::memset(&a, 0, sizeof a);
    for (int h = 0; h < R.G; h++) {
        a.y[h] = reinterpret_cast<const double2 *>(small ? R.ys[seq & 1][h] : R.y[seq & 1][h]);
        a.r[h] = reinterpret_cast<double2 *>(small ? R.rs[h] : R.r[h]);
        a.parts[h] = R.parts[h];
        a.flags[h] = R.flags[h];
    }
    a.G = R.G; a.g = R.g; a.seq = seq;
    a.ctr_target = R.two_shot_calls * (unsigned long long)kRingCtas;
}

static void layout(PeerRing &R, int h, void *base) {
    char *p = static_cast<char *>(base);
    const size_t vec = align_up(R.cap * sizeof(double), 256);
    R.y[0][h] = reinterpret_cast<double *>(p);
    R.y[1][h] = reinterpret_cast<double *>(p + vec);
    R.r[h] = reinterpret_cast<double *>(p + 2 * vec);
    R.parts[h] = reinterpret_cast<double *>(p + 3 * vec);
    char *q = p + 3 * vec + align_up((size_t)kRingMaxRanks * kRingCtas * sizeof(double), 256);
    R.flags[h] = reinterpret_cast<unsigned long long *>(q);
    q += align_up(kFlagWords * sizeof(unsigned long long), 256);
    const size_t sv = align_up(kRingSmallCap * sizeof(double), 256);
    R.ys[0][h] = reinterpret_cast<double *>(q);
    R.ys[1][h] = reinterpret_cast<double *>(q + sv);
    R.rs[h] = reinterpret_cast<double *>(q + 2 * sv);
}

static size_t block_bytes(u64 cap) {
    return 3 * align_up(cap * sizeof(double), 256) + align_up((size_t)kRingMaxRanks * kRingCtas * sizeof(double), 256) +
           align_up(kFlagWords * sizeof(unsigned long long), 256) + 3 * align_up(kRingSmallCap * sizeof(double), 256) + 256;
}

PeerRing *PeerRing::create(Comm *comm, u64 n_doubles) {
    if (!comm || comm->world() < 2) return nullptr;
    if (comm->world() > kRingMaxRanks) throw Error(SVDLAS2_ERR_COMM, "peer ring: more than 8 ranks");
    std::unique_ptr<PeerRing> R(new PeerRing());
    R->G = comm->world(); R->g = comm->rank(); R->cap = n_doubles;
    if (const char *e = std::getenv("SVDLAS2_RING_TWO_SHOT_KB")) R->two_shot_min_bytes = (u64)std::atoll(e) * 1024;
    const size_t bytes = block_bytes(n_doubles);
    // (a failure on one rank must not leave the others waiting in the exchange below: it becomes part of the agreement)
    cudaError_t e = cudaMalloc(&R->base, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); device_pool_trim(); e = cudaMalloc(&R->base, bytes); } // parked / pooled blocks in the way
    if (e == cudaSuccess) e = cudaMemset(R->base, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof mine);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine, R->base);
    else { cudaGetLastError(); if (R->base) { cudaFree(R->base); R->base = nullptr; } }
    // every rank takes part in the exchange even when its own handle could not be made: the ranks must agree on the outcome
    std::vector<unsigned char> all((size_t)R->G * (sizeof mine + 8), 0), me(sizeof mine + 8, 0);
    std::memcpy(me.data(), &mine, sizeof mine);
    me[sizeof mine] = (e == cudaSuccess) ? 1 : 0;
    if (e != cudaSuccess) cudaGetLastError();
    comm->allgather_host_bytes(me.data(), all.data(), me.size());
    bool ok = true;
    for (int h = 0; h < R->G; h++) ok = ok && all[(size_t)h * me.size() + sizeof mine] == 1;
    if (ok) {
        for (int h = 0; h < R->G && ok; h++) {
            if (h == R->g) { R->peer_base[h] = R->base; continue; }
            cudaIpcMemHandle_t hh;
            std::memcpy(&hh, all.data() + (size_t)h * me.size(), sizeof hh);
            e = cudaIpcOpenMemHandle(&R->peer_base[h], hh, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); R->peer_base[h] = nullptr; ok = false; }
        }
    }
    // second round: did everybody manage to map everybody?
    unsigned char mine_ok = ok ? 1 : 0;
    std::vector<unsigned char> oks(R->G * 8, 0), mo(8, 0);
    mo[0] = mine_ok;
    comm->allgather_host_bytes(mo.data(), oks.data(), 8);
    for (int h = 0; h < R->G; h++) ok = ok && oks[(size_t)h * 8] == 1;
    if (!ok) throw Error(SVDLAS2_ERR_COMM, "peer ring: CUDA IPC peer mapping is not available between the ranks");
    for (int h = 0; h < R->G; h++) layout(*R, h, R->peer_base[h]);
    return R.release();
}

void PeerRing::reset_vectors(cudaStream_t s) {
    if (!cap) return;
    LAS2_CUDA(cudaMemsetAsync(y[0][g], 0, cap * sizeof(double), s));
    LAS2_CUDA(cudaMemsetAsync(y[1][g], 0, cap * sizeof(double), s));
    LAS2_CUDA(cudaMemsetAsync(r[g], 0, cap * sizeof(double), s));
}

void PeerRing::close_peers() {
    for (int h = 0; h < G; h++)
        if (h != g && peer_base[h]) { cudaIpcCloseMemHandle(peer_base[h]); peer_base[h] = nullptr; }
}

PeerRing::~PeerRing() {
    close_peers();
    if (base) cudaFree(base);
}

void PeerRing::reduce(u64 ld, const double *qprev, double rnm, const double *q, PartialSlot *dot, cudaStream_t s) {
    if (ld > cap) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "peer ring: vector longer than the ring's buffers");
    seq += 1;
    calls += 1;
    const bool two_shot = ld * sizeof(double) > two_shot_min_bytes;
    if (two_shot) two_shot_calls += 1;
    RingArgs a;
    fill_args(*this, a, seq);
    const int want = (q && dot) ? 1 : 0;
    if (two_shot)
        k_allreduce_epilogue<true, false><<<dim3(kRingCtas, 1), kRingThreads, 0, s>>>(a, nullptr, ld / 2, (const double2 *)qprev, rnm, (const double2 *)q, want);
    else
        k_allreduce_epilogue<false, false><<<dim3(kRingCtas, 1), kRingThreads, 0, s>>>(a, nullptr, ld / 2, (const double2 *)qprev, rnm, (const double2 *)q, want);
    LAS2_LAUNCH_CHECK();
    if (want) {
        dot->parts = parts[g];
        dot->nparts = two_shot ? G * kRingCtas : kRingCtas;
    }
}

void PeerRing::allreduce_small(double *buf, u64 n, cudaStream_t s) {
    if (n == 0) return;
    if (n > kRingSmallCap) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "peer ring: small all-reduce beyond its capacity");
    seq += 1;
    calls += 1;
    const u64 n_even = (n + 1) & ~(u64)1; // the kernel works on double2; the pad element is zero on every rank
    double *mine = ys[seq & 1][g];
    if (n_even != n) LAS2_CUDA(cudaMemsetAsync(mine + n, 0, sizeof(double), s));
    LAS2_CUDA(cudaMemcpyAsync(mine, buf, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    RingArgs a;
    fill_args(*this, a, seq, true);
    const unsigned ctas = (unsigned)std::max<u64>(1, std::min<u64>(16, (n_even / 2 + kRingThreads - 1) / kRingThreads));
    k_allreduce_epilogue<false, false><<<dim3(ctas, 1), kRingThreads, 0, s>>>(a, nullptr, n_even / 2, nullptr, 0.0, nullptr, 0);
    LAS2_LAUNCH_CHECK();
    LAS2_CUDA(cudaMemcpyAsync(buf, rs[g], n * sizeof(double), cudaMemcpyDeviceToDevice, s));
}

// ---- single-GPU emulation of G ranks (tests): one launch, blockIdx.y = rank, all CTAs co-resident ----------------------
// ys: G partial vectors (ld doubles each, contiguous); out: G result vectors; dots: G canonical sums (as a consumer would
// form them).  Returns through the device buffers; used by svdlas2_test_ring_emulation.
void ring_emulate(int G, u64 ld, const double *ys, const double *qprev, double rnm, const double *q, int two_shot, int ctas_per_rank,
                  double *out, double *dots_parts /* G * (G * ctas) */, int *nparts, cudaStream_t s) {
    if (G < 1 || G > kRingMaxRanks || ctas_per_rank < 1 || G * ctas_per_rank > kNumSMs)
        throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "ring emulation: all CTAs must be co-resident (G * ctas <= 148)");
    DevBuf<unsigned long long> flags((size_t)G * kFlagWords);
    LAS2_CUDA(cudaMemsetAsync(flags.get(), 0, (size_t)G * kFlagWords * sizeof(unsigned long long), s));
    DevBuf<RingArgs> dargs(G);
    std::vector<RingArgs> h(G);
    for (int g = 0; g < G; g++) {
        std::memset(&h[g], 0, sizeof(RingArgs));
        for (int k = 0; k < G; k++) {
            h[g].y[k] = reinterpret_cast<const double2 *>(ys + (u64)k * ld);
            h[g].r[k] = reinterpret_cast<double2 *>(out + (u64)k * ld);
            h[g].parts[k] = dots_parts + (size_t)k * G * ctas_per_rank;
            h[g].flags[k] = flags.get() + (size_t)k * kFlagWords;
        }
        h[g].G = G; h[g].g = g; h[g].seq = 1; h[g].ctr_target = (unsigned long long)ctas_per_rank;
    }
    LAS2_CUDA(cudaMemcpyAsync(dargs.get(), h.data(), G * sizeof(RingArgs), cudaMemcpyHostToDevice, s));
    const int want = q ? 1 : 0;
    if (two_shot)
        k_allreduce_epilogue<true, true><<<dim3(ctas_per_rank, G), kRingThreads, 0, s>>>(h[0], dargs.get(), ld / 2, (const double2 *)qprev, rnm, (const double2 *)q, want);
    else
        k_allreduce_epilogue<false, true><<<dim3(ctas_per_rank, G), kRingThreads, 0, s>>>(h[0], dargs.get(), ld / 2, (const double2 *)qprev, rnm, (const double2 *)q, want);
    LAS2_LAUNCH_CHECK();
    k_ring_quiesce<true><<<dim3(1, G), 32, 0, s>>>(h[0], dargs.get());
    LAS2_LAUNCH_CHECK();
    *nparts = two_shot ? G * ctas_per_rank : ctas_per_rank;
    LAS2_CUDA(cudaStreamSynchronize(s));
}

void ring_quiesce(PeerRing &R, cudaStream_t s) {
    R.seq += 1;
    RingArgs a;
    fill_args(R, a, R.seq);
    k_ring_quiesce<false><<<dim3(1, 1), 32, 0, s>>>(a, nullptr);
    LAS2_LAUNCH_CHECK();
}

} // namespace las2
