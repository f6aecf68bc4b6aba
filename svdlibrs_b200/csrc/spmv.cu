// spmv.cu -- y = A x for a device CSR matrix: the two halves of svd_opb (reference src/lib.rs:829-837,
// gather forms :2051-2055 / :2087-2091; the scatter forms :2057-2061 / :2093-2097 never run on the device
// because upload stores both B and B' in row-gather form).
//
// This file holds the dispatch and the kernels that work on the NATURAL CSR arrays: k_spmv_stream for matrices too
// small to fill a persistent kernel (< 2.4 M nnz) and k_spmv_hot as the fallback for large matrices outside the
// index range of the warp-chunk stream.  Large matrices run chunk_spmv.cu (k_spmv_chunk).
//
// Kernel design ("nnz-tile stream"): the value / index arrays are cut into fixed tiles of kSpmvTile
// non-zeros, one CTA per tile, independent of the row structure -- so the 12 B/nnz stream is read with
// perfectly coalesced 128-bit / 64-bit loads even for power-law rows (length 1 .. 1e5+).  A tile
//   1. streams its values and column indices (L1 no-allocate, L2 evict-first), gathers x (L1-cached, L2
//      evict-last) and parks the products in shared memory;
//   2. reduces the products per row: rows of <= 32 entries by one thread each, longer rows by one warp each
//      (lanes stride, shuffle tree) -- both fixed-order, hence deterministic;
//   3. writes y for rows that lie wholly inside the tile; the piece of a row that started in an earlier tile
//      goes to head[tile], the piece of a row that continues into the next tile to carry[tile].
// k_spmv_fixup then stitches the rows that straddle tiles (sum of carries in tile order + head).  No atomics.
//
// Algorithmic bytes per launch: 12*Z + 4*(R+1) + 8*C + 8*R.
#include "spmv.cuh"

#include <algorithm>

namespace las2 {

namespace {

constexpr int kShortRow = 32;  // rows up to this many entries inside a tile are summed by one thread
constexpr int kRowChunk = 1024; // row offsets staged per pass of the reduce phase

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_spmv_stream(const u32 *__restrict__ off, const u32 *__restrict__ idx, const double *__restrict__ val,
              const u32 *__restrict__ tile_row, const double *__restrict__ x, double *__restrict__ y,
              double *__restrict__ carry, double *__restrict__ head, u32 rows, u32 nnz) {
    constexpr int NPT = kSpmvTile / THREADS; // non-zeros per thread
    static_assert(NPT % 2 == 0, "pairs");
    __shared__ __align__(16) double prod[kSpmvTile];
    __shared__ u32 roff[kRowChunk + 1];
    __shared__ u32 longq[kSpmvTile / kShortRow + 2];
    __shared__ u32 longn;

    const u32 tile = blockIdx.x;
    const u32 start = tile * kSpmvTile;
    const u32 end = min(start + kSpmvTile, nnz);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;

    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();

    // ---- phase 1: stream + gather (arrays are zero-padded past nnz, so no bounds checks) ------------
    {
        double2 v[NPT / 2];
        uint2 c[NPT / 2];
#pragma unroll
        for (int k = 0; k < NPT / 2; k++) {
            u32 e = start + 2 * (tid + THREADS * k);
            v[k] = ld_stream_f64x2(val + e, pol_stream);
            c[k] = ld_stream_u32x2(idx + e, pol_stream);
        }
        double2 xv[NPT / 2];
#pragma unroll
        for (int k = 0; k < NPT / 2; k++) {
            xv[k].x = ld_gather_f64(x + c[k].x, pol_keep);
            xv[k].y = ld_gather_f64(x + c[k].y, pol_keep);
        }
#pragma unroll
        for (int k = 0; k < NPT / 2; k++) {
            double2 p;
            p.x = v[k].x * xv[k].x;
            p.y = v[k].y * xv[k].y;
            *reinterpret_cast<double2 *>(&prod[2 * (tid + THREADS * k)]) = p;
        }
    }

    // ---- phase 2/3: per-row reduction ------------------------------------------------------------------
    const u32 R0 = tile_row[tile];
    const u32 R1 = tile_row[tile + 1];           // row that continues past the tile (or `rows`)
    const u32 Rlast = (R1 < rows) ? R1 : rows - 1; // last row with a piece (possibly empty) in this tile
    if (rows == 0) return;

    for (u32 rbase = R0; rbase <= Rlast; rbase += kRowChunk) {
        const u32 nr = min((u32)kRowChunk, Rlast - rbase + 1);
        for (u32 i = tid; i <= nr; i += THREADS) roff[i] = off[rbase + i];
        if (tid == 0) longn = 0;
        __syncthreads(); // also orders phase 1's shared-memory stores before the reads below

        for (u32 rr = tid; rr < nr; rr += THREADS) {
            const u32 r = rbase + rr;
            const u32 o0 = roff[rr], o1 = roff[rr + 1];
            const u32 lo = max(o0, start) - start;
            const u32 hi = min(o1, end) - start;
            if (hi > lo + kShortRow) {
                longq[atomicAdd(&longn, 1u)] = rr; // queue order is irrelevant to the per-row result
            } else {
                double s = 0.0;
                for (u32 k = lo; k < hi; k++) s += prod[k];
                if (r == R1) carry[tile] = s;
                else if (o0 < start) head[tile] = s;
                else y[r] = s;
            }
        }
        __syncthreads();
        const u32 nlong = longn;
        for (u32 qi = warp; qi < nlong; qi += NW) {
            const u32 rr = longq[qi];
            const u32 r = rbase + rr;
            const u32 o0 = roff[rr], o1 = roff[rr + 1];
            const u32 lo = max(o0, start) - start;
            const u32 hi = min(o1, end) - start;
            double s = 0.0;
            for (u32 k = lo + lane; k < hi; k += 32) s += prod[k];
            s = warp_sum(s);
            if (lane == 0) {
                if (r == R1) carry[tile] = s;
                else if (o0 < start) head[tile] = s;
                else y[r] = s;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant with the hot prefix of the gathered vector in shared memory ("k_spmv_hot").
//
// Why: ncu on the first version (profiles/r1a_spmv_ncu_summary.md) shows DRAM at < 30 % with L1TEX at 84 %:
// every gathered x[col] of a warp falls into its own 128-byte line and the L1TEX tag stage retires about one
// line per cycle per SM, so an uncoalesced gather costs ~1 cycle per non-zero per SM no matter how fast HBM is.
// Shared memory serves 32 random 8-byte reads in ~5 cycles.  So:
//   * one CTA per SM (grid = 148), 4 groups of 256 threads; every group walks its own 2048-nnz tiles and
//     synchronises on its own named barrier, so the groups of an SM sit in different phases (stream / reduce);
//   * the first `hot` entries of x (the popular columns of an LSA-shaped matrix: term ids sorted by frequency)
//     are copied to shared memory once per CTA; a gather with col < hot is an LDS, the rest go through L1TEX;
//   * lanes take CONSECUTIVE non-zeros (scalar 8-byte / 4-byte streaming loads), so the 32 gathers of a warp
//     instruction cover 32 adjacent entries of a row and dense regions coalesce into a few lines.
constexpr int kHotGroups = 4;
constexpr int kHotGroupThreads = 256;
constexpr int kHotThreads = kHotGroups * kHotGroupThreads;
constexpr int kHotRowChunk = 1024;

struct HotGroupSmem {
    double prod[kSpmvTile];
    u32 roff[kHotRowChunk + 1];
    u32 longq[kSpmvTile / kShortRow + 2];
    u32 longn;
};

__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kHotGroupThreads) : "memory");
}

__global__ void __launch_bounds__(kHotThreads, 1)
k_spmv_hot(const u32 *__restrict__ off, const u32 *__restrict__ idx, const double *__restrict__ val,
           const u32 *__restrict__ tile_row, const double *__restrict__ x, double *__restrict__ y,
           double *__restrict__ carry, double *__restrict__ head, u32 rows, u32 nnz, u32 num_tiles, u32 hot) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HotGroupSmem *groups = reinterpret_cast<HotGroupSmem *>(smem_raw);
    double *xs = reinterpret_cast<double *>(smem_raw + sizeof(HotGroupSmem) * kHotGroups);

    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();
    for (u32 i = threadIdx.x; i < hot; i += kHotThreads) xs[i] = ld_gather_f64(x + i, pol_keep);
    __syncthreads();

    const int g = threadIdx.x / kHotGroupThreads;
    const int tg = threadIdx.x % kHotGroupThreads;
    const int lane = tg & 31, warp = tg >> 5;
    constexpr int NW = kHotGroupThreads / 32;
    constexpr int NPT = kSpmvTile / kHotGroupThreads;
    HotGroupSmem &S = groups[g];

    for (u32 tile = blockIdx.x * kHotGroups + g; tile < num_tiles; tile += gridDim.x * kHotGroups) {
        const u32 start = tile * kSpmvTile;
        const u32 end = min(start + kSpmvTile, nnz);
        const u32 R0 = tile_row[tile];
        const u32 R1 = tile_row[tile + 1];
        // ---- stream + gather: lane-consecutive non-zeros ----
        {
            double v[NPT];
            u32 c[NPT];
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                const u32 e = start + k * kHotGroupThreads + tg;
                v[k] = ld_stream_f64(val + e, pol_stream);
                c[k] = ld_stream_u32(idx + e, pol_stream);
            }
            double xv[NPT];
#pragma unroll
            for (int k = 0; k < NPT; k++) {
                if (c[k] < hot) xv[k] = xs[c[k]];
                else xv[k] = ld_gather_f64(x + c[k], pol_keep);
            }
#pragma unroll
            for (int k = 0; k < NPT; k++) S.prod[k * kHotGroupThreads + tg] = v[k] * xv[k];
        }
        // ---- per-row reduction (same scheme as k_spmv_stream) ----
        const u32 Rlast = (R1 < rows) ? R1 : rows - 1;
        for (u32 rbase = R0; rbase <= Rlast; rbase += kHotRowChunk) {
            const u32 nr = min((u32)kHotRowChunk, Rlast - rbase + 1);
            for (u32 i = tg; i <= nr; i += kHotGroupThreads) S.roff[i] = off[rbase + i];
            if (tg == 0) S.longn = 0;
            group_sync(g);
            for (u32 rr = tg; rr < nr; rr += kHotGroupThreads) {
                const u32 r = rbase + rr;
                const u32 o0 = S.roff[rr], o1 = S.roff[rr + 1];
                const u32 lo = max(o0, start) - start;
                const u32 hi = min(o1, end) - start;
                if (hi > lo + kShortRow) {
                    S.longq[atomicAdd(&S.longn, 1u)] = rr;
                } else {
                    double s = 0.0;
                    for (u32 k = lo; k < hi; k++) s += S.prod[k];
                    if (r == R1) carry[tile] = s;
                    else if (o0 < start) head[tile] = s;
                    else y[r] = s;
                }
            }
            group_sync(g);
            const u32 nlong = S.longn;
            for (u32 qi = warp; qi < nlong; qi += NW) {
                const u32 rr = S.longq[qi];
                const u32 r = rbase + rr;
                const u32 o0 = S.roff[rr], o1 = S.roff[rr + 1];
                const u32 lo = max(o0, start) - start;
                const u32 hi = min(o1, end) - start;
                double s = 0.0;
                for (u32 k = lo + lane; k < hi; k += 32) s += S.prod[k];
                s = warp_sum(s);
                if (lane == 0) {
                    if (r == R1) carry[tile] = s;
                    else if (o0 < start) head[tile] = s;
                    else y[r] = s;
                }
            }
            group_sync(g);
        }
    }
}

// Rows that straddle tile boundaries: y[row] = carry[first] + ... + carry[b-1] + head[b], b = the tile the row
// ends in.  (row, first) per tile are precomputed at upload (k_fix_list), so this is two dependent load levels.
__global__ void k_spmv_fixup(const u32 *__restrict__ fix_row, const u32 *__restrict__ fix_first,
                             const double *__restrict__ carry, const double *__restrict__ head,
                             double *__restrict__ y, u32 num_tiles) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (b >= num_tiles) return;
    const u32 row = fix_row[b];
    if (row == 0xffffffffu) return;
    const u32 first = fix_first[b];
    const double h = head[b];
    double s = 0.0;
    for (u32 t = first; t < b; t++) s += carry[t];
    y[row] = s + h;
}

} // namespace


namespace {
constexpr u32 kHotMaxEntries = 16384; // upper bound of the knob: 128 KB of the 227 KB
size_t hot_smem_bytes(u32 hot) { return sizeof(HotGroupSmem) * kHotGroups + (size_t)hot * sizeof(double); }
bool g_hot_attr_set[kMaxDevicesTracked]; // per device
} // namespace

void spmv_chunk(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches); // chunk_spmv.cu

void spmv(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches) {
    if (A.rows == 0) return;
    if (A.chk.active && (A.tune.variant < 0 || A.tune.variant == 5)) {
        spmv_chunk(A, x, y, s, launches);
        return;
    }
    // variant: -1 auto, 0/1/2 = one CTA per tile with 256/128/512 threads, 3 = persistent hot-prefix kernel
    int variant = A.tune.variant;
    if (variant < 0) variant = (A.num_tiles >= 8u * kNumSMs) ? 3 : 0;
    switch (variant) {
    case 3: {
        u32 hot = A.tune.hot_entries >= 0 ? (u32)A.tune.hot_entries : A.hot_entries;
        if (hot > kHotMaxEntries) hot = kHotMaxEntries;
        if (hot > A.cols) hot = (u32)A.cols;
        bool &attr_set = g_hot_attr_set[current_device_slot()];
        if (!attr_set) {
            LAS2_CUDA(cudaFuncSetAttribute(k_spmv_hot, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)hot_smem_bytes(kHotMaxEntries)));
            attr_set = true;
        }
        u32 grid = kNumSMs;
        if (grid * kHotGroups > A.num_tiles) grid = cdiv(A.num_tiles, kHotGroups);
        k_spmv_hot<<<grid, kHotThreads, hot_smem_bytes(hot), s>>>(A.off, A.idx, A.val, A.tile_row, x, y, A.carry, A.head,
                                                                 (u32)A.rows, (u32)A.nnz, A.num_tiles, hot);
        break;
    }
    case 1:
        k_spmv_stream<128><<<A.num_tiles, 128, 0, s>>>(A.off, A.idx, A.val, A.tile_row, x, y, A.carry, A.head,
                                                       (u32)A.rows, (u32)A.nnz);
        break;
    case 2:
        k_spmv_stream<512><<<A.num_tiles, 512, 0, s>>>(A.off, A.idx, A.val, A.tile_row, x, y, A.carry, A.head,
                                                       (u32)A.rows, (u32)A.nnz);
        break;
    default:
        k_spmv_stream<256><<<A.num_tiles, 256, 0, s>>>(A.off, A.idx, A.val, A.tile_row, x, y, A.carry, A.head,
                                                       (u32)A.rows, (u32)A.nnz);
        break;
    }
    if (A.num_tiles > 1) {
        k_spmv_fixup<<<cdiv(A.num_tiles - 1, 128), 128, 0, s>>>(A.fix_row, A.fix_first, A.carry, A.head, y, A.num_tiles);
        if (launches) *launches += 1;
    }
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 1;
}

} // namespace las2
