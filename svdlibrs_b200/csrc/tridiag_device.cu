// tridiag_device.cu -- the O(js^3) half of imtql2 on the GPU.
//
// The reference updates all js rows of the eigenvector matrix z for every Givens rotation of the implicit-QL
// iteration (src/lib.rs:1647-1652, stride-js inner loop: ~0.5 s at js=800, minutes at js=3000 on one core).
// The rotation parameters depend only on the tridiagonal (d, e), never on z, so the host runs the cheap
// O(js^2) recurrences (host_math.cpp: ql_rotations) and this kernel replays the recorded rotations:
// one thread per row of z, rows independent, z stored transposed so that the threads of a warp touch
// consecutive addresses.  Within a sweep, rotation i touches columns (i, i+1) and i descends, so column i+1
// is final after rotation i and column i is carried in a register: one load + one store per rotation per
// row (16 B), the minimum.  Arithmetic is the reference's: two rounded products and one rounded add/sub.
#include "tridiag_device.cuh"

#include <cstdlib>

namespace las2 {

namespace {

__global__ void k_identity(double *__restrict__ zT, u32 ldz, u32 js) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 total = (u64)ldz * js;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        u32 col = (u32)(e / ldz), row = (u32)(e % ldz);
        zT[e] = (row == col) ? 1.0 : 0.0;
    }
}

constexpr int kReplayThreads = 64;

__global__ void __launch_bounds__(kReplayThreads)
k_ql_replay(double *__restrict__ zT, u32 ldz, u32 js, const QlSweep *__restrict__ sweeps, u32 nsweeps,
            const double2 *__restrict__ sc) {
    const u32 row = blockIdx.x * kReplayThreads + threadIdx.x;
    if (row >= js) return;
    double *z = zT + row;
    for (u32 w = 0; w < nsweeps; w++) {
        const QlSweep sw = sweeps[w];
        const double2 *__restrict__ rot = sc + sw.first;
        u32 i = sw.hi;
        double carry = z[(u64)(i + 1) * ldz];
#pragma unroll 4
        for (u32 t = 0; t < sw.count; t++, i--) {
            const double2 p = rot[t];                  // (s, c)
            const double x = z[(u64)i * ldz];
            z[(u64)(i + 1) * ldz] = __dadd_rn(__dmul_rn(p.x, x), __dmul_rn(p.y, carry));
            carry = __dsub_rn(__dmul_rn(p.y, x), __dmul_rn(p.x, carry));
        }
        z[(u64)(i + 1) * ldz] = carry; // i has been decremented once past the last rotation index
    }
}

// Same replay with the CTA's rows of z resident in shared memory (js x rpc doubles): the per-rotation load is an
// LDS instead of an L2 round trip, which matters because the rotations of one row form a dependent chain.
__global__ void k_ql_replay_smem(double *__restrict__ zT, u32 ldz, u32 js, u32 rpc, const QlSweep *__restrict__ sweeps,
                                 u32 nsweeps, const double2 *__restrict__ sc) {
    extern __shared__ double zs[]; // zs[col * rpc + local_row]
    const u32 row0 = blockIdx.x * rpc;
    for (u32 e = threadIdx.x; e < js * rpc; e += blockDim.x) {
        const u32 col = e / rpc, lr = e % rpc;
        zs[e] = (row0 + lr == col) ? 1.0 : 0.0;
    }
    __syncthreads();
    // One warp replays: lane = local row.  The rotation parameters are the same for every row, so the warp
    // fetches them 32 at a time (one coalesced 512-byte load, prefetched one batch ahead) and broadcasts them
    // with shuffles instead of stalling on a uniform global load per rotation.
    if (threadIdx.x < 32) {
        const u32 lane = threadIdx.x;
        const bool live = lane < rpc && row0 + lane < js;
        double *z = zs + (live ? lane : 0);
        for (u32 w = 0; w < nsweeps; w++) {
            const QlSweep sw = sweeps[w];
            const double2 *__restrict__ rot = sc + sw.first;
            u32 i = sw.hi;
            double carry = z[(i + 1) * rpc];
            double2 nxt = (lane < sw.count) ? rot[lane] : make_double2(0.0, 0.0);
            for (u32 t0 = 0; t0 < sw.count; t0 += 32) {
                const double2 cur = nxt;
                if (t0 + 32 + lane < sw.count) nxt = rot[t0 + 32 + lane];
                const u32 nb = min(32u, sw.count - t0);
#pragma unroll 4
                for (u32 t = 0; t < nb; t++, i--) {
                    const double ps = __shfl_sync(0xffffffffu, cur.x, t);
                    const double pc = __shfl_sync(0xffffffffu, cur.y, t);
                    const double x = z[i * rpc];
                    const double hi_new = __dadd_rn(__dmul_rn(ps, x), __dmul_rn(pc, carry));
                    carry = __dsub_rn(__dmul_rn(pc, x), __dmul_rn(ps, carry));
                    if (live) z[(i + 1) * rpc] = hi_new;
                }
            }
            if (live) z[(u32)(i + 1) * rpc] = carry;
        }
    }
    __syncthreads();
    for (u32 e = threadIdx.x; e < js * rpc; e += blockDim.x) {
        const u32 col = e / rpc, r = e % rpc;
        if (row0 + r < js) zT[(u64)col * ldz + row0 + r] = zs[e];
    }
}

// ---- the same replay as a PIPELINE OF SWEEPS over the warps of the CTA ----------------------------------------------
// The rotations of one row form a dependent chain, so a single warp per group of rows is bound by the latency of that
// chain (~90 clocks per rotation, profiles/r2v_ql_pipeline.md) while the rows themselves are few (js <= a few thousand).
// But sweep v+1 only needs the columns sweep v has already left behind: rotation i of a sweep touches columns (i, i+1)
// and the indices descend, so once sweep v has finished rotation k its columns > k are final as far as v is concerned.
// Warp w therefore replays sweeps w, w + W, w + 2W, ... and follows its predecessor (the warp replaying sweep v - 1) at a
// distance of one batch of 32 rotations -- the classic wavefront over a sequence of Givens sweeps.  Every element of z
// still sees exactly the same operations in exactly the same order: the result is bit-identical to the one-warp replay.
//
// Protocol (one 64-bit word per warp in shared memory: sweep id << 32 | progress):
//   * progress p of a running sweep = every column >= p is final as far as this sweep AND all earlier ones are
//     concerned; it is published after each batch (p = lowest finished rotation index + 1) and is "infinite" until the
//     first batch is done, so that a follower can never overtake a predecessor of its predecessor;
//   * a sweep retires (its warp publishes the id of its NEXT sweep) only after its predecessor has retired: retirement
//     is in order, hence "id in the predecessor's slot > v - 1" means that every earlier sweep is complete;
//   * before a batch whose lowest rotation index is i_low a warp waits for "predecessor retired or its p <= i_low".
// A waiting warp only ever waits for an older sweep, and the oldest one waits for nobody: no deadlock.
constexpr int kReplayWarps = 8;
constexpr u32 kProgressInf = 0x7fffffffu;
constexpr size_t kReplayPipeHeader = 128 + (size_t)kReplayWarps * 2 * 32 * sizeof(double2); // slots + rotation batches

__global__ void __launch_bounds__(kReplayWarps * 32)
k_ql_replay_pipe(double *__restrict__ zT, u32 ldz, u32 js, u32 rpc, const QlSweep *__restrict__ sweeps, u32 nsweeps,
                 const double2 *__restrict__ sc) {
    extern __shared__ __align__(16) unsigned char ql_smem[];
    volatile unsigned long long *slot = reinterpret_cast<volatile unsigned long long *>(ql_smem);
    double2 *rotbuf = reinterpret_cast<double2 *>(ql_smem + 128);
    double *zs = reinterpret_cast<double *>(ql_smem + kReplayPipeHeader); // zs[col * rpc + local_row]
    constexpr u32 W = kReplayWarps;
    const u32 row0 = blockIdx.x * rpc;
    for (u32 e = threadIdx.x; e < js * rpc; e += blockDim.x) {
        const u32 col = e / rpc, lr = e % rpc;
        zs[e] = (row0 + lr == col) ? 1.0 : 0.0;
    }
    if (threadIdx.x < W) slot[threadIdx.x] = ((unsigned long long)threadIdx.x << 32) | kProgressInf;
    __syncthreads();

    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool live = lane < rpc && row0 + lane < js;
    double *z = zs + (live ? lane : 0);
    const u32 pw = (warp + W - 1) % W; // the warp that replays the preceding sweep
    for (u32 v = warp; v < nsweeps; v += W) {
        const QlSweep sw = sweeps[v];
        const double2 *__restrict__ rot = sc + sw.first;
        bool pred_retired = (v == 0);
        u32 i = sw.hi;
        double carry = 0.0;
        double2 nxt = (lane < sw.count) ? rot[lane] : make_double2(0.0, 0.0);
        for (u32 t0 = 0; t0 < sw.count; t0 += 32) {
            const u32 nb = min(32u, sw.count - t0);
            const u32 i_low = i - (nb - 1); // lowest rotation index of this batch
            double2 *rb = rotbuf + (warp * 2 + ((t0 >> 5) & 1u)) * 32;
            rb[lane] = nxt; // the batch's (s, c) pairs: read back below as one broadcast LDS.128 per rotation
            if (t0 + 32 + lane < sw.count) nxt = rot[t0 + 32 + lane];
            if (!pred_retired) {
                for (;;) {
                    const unsigned long long sv = slot[pw];
                    if ((u32)(sv >> 32) > v - 1) { pred_retired = true; break; }
                    if ((u32)sv <= i_low) break;
                }
            }
            __threadfence_block(); // what the predecessor stored before publishing is visible from here on
            __syncwarp();
            if (t0 == 0) carry = z[(i + 1) * rpc];
#pragma unroll 4
            for (u32 t = 0; t < nb; t++, i--) {
                const double2 p = rb[t]; // (s, c)
                const double x = z[i * rpc];
                const double hi_new = __dadd_rn(__dmul_rn(p.x, x), __dmul_rn(p.y, carry));
                carry = __dsub_rn(__dmul_rn(p.y, x), __dmul_rn(p.x, carry));
                if (live) z[(i + 1) * rpc] = hi_new;
            }
            if (t0 + 32 < sw.count) { // rotations down to index i + 1 are done: columns >= i + 2 are final
                __threadfence_block();
                __syncwarp();
                if (lane == 0) slot[warp] = ((unsigned long long)v << 32) | (unsigned long long)(i + 2);
            }
        }
        if (sw.count && live) z[(u32)(i + 1) * rpc] = carry; // i has been decremented once past the last rotation index
        if (!pred_retired) {
            while ((u32)(slot[pw] >> 32) <= v - 1) {}
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) slot[warp] = ((unsigned long long)(v + W) << 32) | kProgressInf; // retire: next sweep, nothing final yet
    }
    __syncthreads();
    for (u32 e = threadIdx.x; e < js * rpc; e += blockDim.x) {
        const u32 col = e / rpc, r = e % rpc;
        if (row0 + r < js) zT[(u64)col * ldz + row0 + r] = zs[e];
    }
}

__global__ void k_ritz_coefficients(const double *__restrict__ zT, u32 ldz, u32 js, const u32 *__restrict__ phys,
                                    u32 d, double *__restrict__ C) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (u64)js * d) return;
    u32 c = (u32)(e / js), i = (u32)(e % js); // consecutive threads walk down one column of z: coalesced reads
    C[(u64)i * d + c] = zT[(u64)phys[c] * ldz + (js - 1 - i)];
}

} // namespace

int g_ql_replay_variant = -1; // test / tuning hook: -1 = default (env SVDLAS2_QL_PIPE, pipeline on), 0 = one warp per CTA, 1 = pipeline

void ql_replay(const QlPlan &plan, u32 js, u32 ldz, double *zT, cudaStream_t s, u64 *launches, u64 *h2d_bytes) {
    static const bool pipe_env = [] { const char *e = std::getenv("SVDLAS2_QL_PIPE"); return !(e && e[0] == '0'); }();
    const bool pipe = g_ql_replay_variant < 0 ? pipe_env : g_ql_replay_variant == 1;
    const size_t header = pipe ? kReplayPipeHeader : 0;
    // rows of z per CTA such that js x rpc doubles fit in shared memory (largest power of two <= 32)
    u32 rpc = 32;
    while (rpc > 4 && (size_t)js * rpc * sizeof(double) + header > 200 * 1024) rpc >>= 1;
    const bool in_smem = (size_t)js * rpc * sizeof(double) + header <= 200 * 1024;
    if (!in_smem || plan.sweeps.empty()) {
        unsigned g = cdiv((u64)ldz * js, 256);
        if (g > (unsigned)kNumSMs * 16) g = kNumSMs * 16;
        k_identity<<<g, 256, 0, s>>>(zT, ldz, js);
        if (launches) *launches += 1;
    }
    if (!plan.sweeps.empty()) {
        DevBuf<QlSweep> dsw(plan.sweeps.size());
        DevBuf<double> dsc(plan.sc.size());
        LAS2_CUDA(cudaMemcpyAsync(dsw.get(), plan.sweeps.data(), plan.sweeps.size() * sizeof(QlSweep),
                                  cudaMemcpyHostToDevice, s));
        LAS2_CUDA(cudaMemcpyAsync(dsc.get(), plan.sc.data(), plan.sc.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        if (h2d_bytes) *h2d_bytes += plan.sweeps.size() * sizeof(QlSweep) + plan.sc.size() * sizeof(double);
        if (in_smem) {
            const size_t bytes = (size_t)js * rpc * sizeof(double) + header;
            static bool attr_set_dev[kMaxDevicesTracked];
            bool &attr_set = attr_set_dev[current_device_slot()];
            if (!attr_set) {
                LAS2_CUDA(cudaFuncSetAttribute(k_ql_replay_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                LAS2_CUDA(cudaFuncSetAttribute(k_ql_replay_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set = true;
            }
            if (pipe)
                k_ql_replay_pipe<<<cdiv(js, rpc), kReplayWarps * 32, bytes, s>>>(zT, ldz, js, rpc, dsw, (u32)plan.sweeps.size(),
                                                                                 reinterpret_cast<const double2 *>(dsc.get()));
            else
                k_ql_replay_smem<<<cdiv(js, rpc), 128, bytes, s>>>(zT, ldz, js, rpc, dsw, (u32)plan.sweeps.size(),
                                                                   reinterpret_cast<const double2 *>(dsc.get()));
        } else {
            k_ql_replay<<<cdiv(js, kReplayThreads), kReplayThreads, 0, s>>>(zT, ldz, js, dsw, (u32)plan.sweeps.size(),
                                                                          reinterpret_cast<const double2 *>(dsc.get()));
        }
        LAS2_LAUNCH_CHECK();
        if (launches) *launches += 1;
        LAS2_CUDA(cudaStreamSynchronize(s)); // dsw / dsc are freed on scope exit
    }
}

void ritz_coefficients(const double *zT, u32 ldz, u32 js, const u32 *phys_cols_dev, u32 d, double *C, cudaStream_t s,
                       u64 *launches) {
    if (d == 0) return;
    k_ritz_coefficients<<<cdiv((u64)js * d, 256), 256, 0, s>>>(zT, ldz, js, phys_cols_dev, d, C);
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 1;
}

} // namespace las2
