// spmv.cu -- y = A x for a device CSR matrix: the two halves of svd_opb (reference src/lib.rs:829-837,
// gather forms :2051-2055 / :2087-2091; the scatter forms :2057-2061 / :2093-2097 never run on the device
// because upload stores both B and B' in row-gather form).
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

// ---------------------------------------------------------------------------------------------------------
// Column-blocked SpMV ("k_spmv_blocked").  Format: BlockedCsr (device_csr.cuh / blocked_csr.cu).
//
// What ncu said about the earlier kernels (profiles/r1a..r1c): DRAM < 30 % busy while the L1TEX / shared-memory
// data pipe ran at 58-84 % -- first from uncoalesced x[col] gathers (one 128-byte line per cycle per SM), then,
// once the gathers came from shared memory, from parking every product in shared memory and re-reading it for
// the per-row sums (0.77 wavefronts and ~70 thread instructions per non-zero).  This kernel therefore
//   * gathers only from shared memory on slab tiles (the 128 KB slice of x of the slab's column block),
//   * streams the non-zeros with TMA bulk copies (cp.async.bulk + mbarrier) into a per-group stage,
//   * keeps the products in REGISTERS: inside a tile the stream is stored interleaved (element t*8+k at slot
//     k*128+t), so the conflict-free read "slot k*128 + lane" hands every lane 8 CONSECUTIVE elements of the
//     row-ordered stream; the lane sums them serially, cutting at the row starts (a 1024-bit map written by
//     one thread per row), and a warp-shuffle segmented scan + a 4-entry cross-warp step stitch the rows that
//     span lanes.  No product ever touches shared memory; every order is fixed by the tile: deterministic.
// One CTA per SM, 5 groups of 128 threads, each with its own stage / mbarrier / named barrier, walking its own
// tiles.  Each CTA takes a contiguous range of the slab tiles (one or two column blocks -> one or two loads of
// the x slice) and an interleaved share of the remainder tiles (32-bit columns, gathered through L1/L2).
// Output: one partial per virtual row (slab, row); k_slab_reduce adds the slabs of a row in slab order.
constexpr int kBlkGroupsSlab = 5;  // beside the 128 KB slice of x
constexpr int kBlkGroupsPlain = 8; // no slabs: only the (<= 64 KB) hot prefix of x is resident
constexpr u32 kPlainHotDefault = 4096; // sweep on cfg 2: 4096 entries beat 8192 (more L1 left for the cold gathers)
constexpr int kBlkLanes = 128; // threads per group
constexpr int kBlkE = kBlkTile / kBlkLanes; // 8 consecutive elements per lane
constexpr int kBlkWarps = kBlkLanes / 32;
constexpr int kBlkRowChunk = 256;
constexpr int kBlkRoffPerThread = (kBlkRowChunk + 1 + kBlkLanes - 1) / kBlkLanes; // 3
static_assert(kBlkE == 8, "the flag byte per lane assumes 8 elements per lane");

struct __align__(128) BlkGroupSmem {
    double sval[kBlkTile];        // TMA stage: values (interleaved order)
    u32 sidx[kBlkTile];           // TMA stage: u16 (slab tiles, first half used) or u32 (remainder) indices
    u32 rowat[kBlkTile];          // virtual row whose piece starts at this element (valid where the flag is set)
    u32 roff[kBlkRowChunk + 1];
    u32 flags[kBlkTile / 32];     // bit e: a row piece starts at element e
    double wsum[kBlkWarps];       // warp aggregates of the segmented scan
    u32 wfr[kBlkWarps];
    unsigned long long full_bar;  // mbarrier: the stage has landed
};

__device__ __forceinline__ void blk_group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kBlkLanes) : "memory");
}
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, unsigned long long *bar, u64 pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr u32 kSegFlag = 0x80000000u; // scan element: (sum, fr) with fr = kSegFlag | row when a piece starts in it

// One group's walk over the tiles first, first+step, ... < stop.  SLAB: u16 local indices + shared-memory gathers;
// otherwise u32 indices + global gathers.
template <bool SLAB>
__device__ __forceinline__ void blk_walk(BlkGroupSmem &S, int g, int t, u32 &parity, u32 first, u32 stop, u32 step,
                                         const u32 *__restrict__ voff, const uint16_t *__restrict__ idx16,
                                         const u32 *__restrict__ idx32, u32 rem_tile0, const double *__restrict__ val,
                                         const u32 *__restrict__ tile_row, const double *__restrict__ x,
                                         const double *xs, u32 hot, double *__restrict__ out, double *__restrict__ carry,
                                         double *__restrict__ head, u32 vrows, u64 pol_stream, u64 pol_keep, bool use_tma) {
    constexpr u32 kIdxBytes = SLAB ? kBlkTile * 2 : kBlkTile * 4;
    const int lane = t & 31, warp = t >> 5;
    auto issue = [&](u32 tile) {
        mbar_expect_tx(&S.full_bar, kBlkTile * 8 + kIdxBytes);
        tma_load_1d(S.sval, val + (u64)tile * kBlkTile, kBlkTile * 8, &S.full_bar, pol_stream);
        if (SLAB) tma_load_1d(S.sidx, idx16 + (u64)tile * kBlkTile, kIdxBytes, &S.full_bar, pol_stream);
        else tma_load_1d(S.sidx, idx32 + (u64)(tile - rem_tile0) * kBlkTile, kIdxBytes, &S.full_bar, pol_stream);
    };
    if (first >= stop) return;
    if (t == 0 && use_tma) issue(first);
    for (u32 tile = first; tile < stop; tile += step) {
        const u32 start = tile * kBlkTile, end = start + kBlkTile;
        const u32 R0 = tile_row[tile], R1 = tile_row[tile + 1];
        const u32 Rlast = (R1 < vrows) ? R1 : vrows - 1;
        // row offsets of the first row chunk: in flight while the stage lands
        u32 ro[kBlkRoffPerThread];
        const u32 nr0 = min((u32)kBlkRowChunk, Rlast - R0 + 1);
#pragma unroll
        for (int i = 0; i < kBlkRoffPerThread; i++) {
            const u32 e = t + i * kBlkLanes;
            ro[i] = (e <= nr0) ? voff[R0 + e] : 0u;
        }
        if (use_tma) {
            mbar_wait(&S.full_bar, parity);
            parity ^= 1u;
        } else { // debugging path: cooperative copy instead of the bulk copy
            for (int e = t; e < (int)kBlkTile; e += kBlkLanes) S.sval[e] = val[(u64)tile * kBlkTile + e];
            if (SLAB) { for (int e = t; e < (int)kBlkTile; e += kBlkLanes) reinterpret_cast<uint16_t *>(S.sidx)[e] = idx16[(u64)tile * kBlkTile + e]; }
            else { for (int e = t; e < (int)kBlkTile; e += kBlkLanes) S.sidx[e] = idx32[(u64)(tile - rem_tile0) * kBlkTile + e]; }
            blk_group_sync(g);
        }
        // ---- stage -> registers: slot k*128 + t holds element t*8 + k ----
        double p[kBlkE];
        u32 c[kBlkE];
#pragma unroll
        for (int k = 0; k < kBlkE; k++) {
            p[k] = S.sval[k * kBlkLanes + t];
            c[k] = SLAB ? (u32) reinterpret_cast<const uint16_t *>(S.sidx)[k * kBlkLanes + t] : S.sidx[k * kBlkLanes + t];
        }
        if (t < kBlkTile / 32) S.flags[t] = 0;
#pragma unroll
        for (int i = 0; i < kBlkRoffPerThread; i++) {
            const u32 e = t + i * kBlkLanes;
            if (e <= nr0) S.roff[e] = ro[i];
        }
        {
            // A warp may ARRIVE at the barrier with its shared-memory loads of the stage still queued behind other
            // warps' uncoalesced gathers (the LSU queue is in-order and a 32-line gather holds it for 32 cycles), and
            // the bulk copy issued right after the barrier would then overwrite the stage under them (seen on B200
            // as a handful of wrong rows per 45 M non-zeros).  A branch on the loaded values forces every stage
            // read to COMPLETE before the barrier; it is never taken.
            u32 chk = 0;
#pragma unroll
            for (int k = 0; k < kBlkE; k++) chk ^= c[k] ^ (u32)__double2loint(p[k]) ^ (u32)__double2hiint(p[k]);
            if (chk == 0x9e3779b9u && c[0] == 0xffffffffu) S.flags[0] = 0xffffffffu;
        }
        blk_group_sync(g); // stage consumed, flags cleared, first row chunk staged
        if (t == 0 && use_tma && tile + step < stop) issue(tile + step);
        // ---- gathers + products (registers only) ----
        if (SLAB) {
#pragma unroll
            for (int k = 0; k < kBlkE; k++) p[k] *= xs[c[k]];
        } else {
            double xv[kBlkE];
#pragma unroll
            for (int k = 0; k < kBlkE; k++) {
                if (c[k] < hot) xv[k] = xs[c[k]];
                else xv[k] = ld_gather_f64(x + c[k], pol_keep);
            }
#pragma unroll
            for (int k = 0; k < kBlkE; k++) p[k] *= xv[k];
        }
        // ---- one thread per row: mark where its piece starts; rows with no entry in this tile are zero ----
        const bool first_is_head = S.roff[0] < start; // row R0 began in an earlier tile
        auto emit = [&](u32 r, double v) {
            if (r == R1) carry[tile] = v;
            else if (r == R0 && first_is_head) head[tile] = v;
            else out[r] = v;
        };
        for (u32 rbase = R0; rbase <= Rlast; rbase += kBlkRowChunk) {
            const u32 nr = min((u32)kBlkRowChunk, Rlast - rbase + 1);
            if (rbase != R0) {
                blk_group_sync(g); // everyone is done with the previous chunk's offsets
                for (u32 i = t; i <= nr; i += kBlkLanes) S.roff[i] = voff[rbase + i];
                blk_group_sync(g);
            }
            for (u32 rr = t; rr < nr; rr += kBlkLanes) {
                const u32 r = rbase + rr;
                const u32 o0 = S.roff[rr], o1 = S.roff[rr + 1];
                const u32 lo = max(o0, start) - start;
                const u32 hi = min(o1, end) - start;
                if (hi > lo) {
                    atomicOr(&S.flags[lo >> 5], 1u << (lo & 31)); // integer bit set: order-independent
                    S.rowat[lo] = r;
                } else {
                    emit(r, 0.0);
                }
            }
        }
        blk_group_sync(g); // flags and rowat complete
        // ---- per-lane serial segmented sum over its 8 consecutive elements ----
        const u32 fl = (S.flags[t >> 2] >> ((t & 3) * 8)) & 0xffu;
        double s = 0.0, head_sum = 0.0;
        bool open_here = false;
        u32 open_row = 0;
#pragma unroll
        for (int k = 0; k < kBlkE; k++) {
            if ((fl >> k) & 1u) {
                if (open_here) emit(open_row, s);
                else head_sum = s;
                s = 0.0;
                open_here = true;
                open_row = S.rowat[t * kBlkE + k];
            }
            s += p[k];
        }
        // ---- stitch pieces that span lanes: warp segmented scan, then the 4 warp aggregates ----
        double xsum = s;
        u32 fr = open_here ? (kSegFlag | open_row) : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ys = __shfl_up_sync(0xffffffffu, xsum, o);
            const u32 yf = __shfl_up_sync(0xffffffffu, fr, o);
            if (lane >= o && !(fr & kSegFlag)) { xsum = ys + xsum; fr = yf; }
        }
        if (lane == 31) { S.wsum[warp] = xsum; S.wfr[warp] = fr; }
        // exclusive value inside the warp
        double cin_sum = __shfl_up_sync(0xffffffffu, xsum, 1);
        u32 cin_fr = __shfl_up_sync(0xffffffffu, fr, 1);
        if (lane == 0) { cin_sum = 0.0; cin_fr = 0u; }
        blk_group_sync(g);
        // fold the aggregates of the warps to the left (in order) in front of it
        double acc_sum = 0.0;
        u32 acc_fr = 0u;
        for (int w = 0; w < warp; w++) {
            const double ws = S.wsum[w];
            const u32 wf = S.wfr[w];
            if (wf & kSegFlag) { acc_sum = ws; acc_fr = wf; }
            else acc_sum = acc_sum + ws;
        }
        if (!(cin_fr & kSegFlag)) { cin_sum = acc_sum + cin_sum; cin_fr = acc_fr; }
        // a lane that starts a piece closes the piece that was open when it began
        if (open_here) {
            if (cin_fr & kSegFlag) emit(cin_fr & ~kSegFlag, cin_sum + head_sum);
        }
        // the piece still open at the end of the tile
        if (t == kBlkLanes - 1) {
            if (open_here) emit(open_row, s);
            else if (cin_fr & kSegFlag) emit(cin_fr & ~kSegFlag, cin_sum + s);
        }
        // (the next tile's first barrier separates these reads of wsum / rowat / flags from their next writes)
    }
}

template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * kBlkLanes, 1)
k_spmv_blocked(const u32 *__restrict__ voff, const uint16_t *__restrict__ idx16, const u32 *__restrict__ idx32,
               const double *__restrict__ val, const u32 *__restrict__ tile_row, const u32 *__restrict__ slab_tile_start,
               const double *__restrict__ x, double *__restrict__ out, double *__restrict__ carry,
               double *__restrict__ head, u32 vrows, u32 cols, u32 nbk, u32 blocked_tiles, u32 num_tiles, u32 hot, int use_tma) {
    constexpr int kBlkGroups = GROUPS;
    constexpr int kBlkThreads = GROUPS * kBlkLanes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlkGroupSmem *groups = reinterpret_cast<BlkGroupSmem *>(smem_raw);
    double *xs = reinterpret_cast<double *>(smem_raw + sizeof(BlkGroupSmem) * kBlkGroups);

    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();
    const int g = threadIdx.x / kBlkLanes;
    const int t = threadIdx.x % kBlkLanes;
    BlkGroupSmem &S = groups[g];
    if (t == 0) {
        mbar_init(&S.full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    u32 parity = 0;
    __syncthreads();

    // ---- slab tiles: contiguous share of [0, blocked_tiles) ----
    const u32 t_begin = (u32)(((u64)blocked_tiles * blockIdx.x) / gridDim.x);
    const u32 t_end = (u32)(((u64)blocked_tiles * (blockIdx.x + 1)) / gridDim.x);
    u32 slab = 0;
    for (u32 tt = t_begin; tt < t_end;) {
        while (slab + 1 < nbk && slab_tile_start[slab + 1] <= tt) slab++;
        const u32 t_stop = min(t_end, slab_tile_start[slab + 1]);
        __syncthreads(); // every group is done with the previous slice of x
        const u32 c0 = slab * kBlockWidth;
        for (u32 i = threadIdx.x; i < kBlockWidth; i += kBlkThreads)
            xs[i] = (c0 + i < cols) ? ld_gather_f64(x + c0 + i, pol_keep) : 0.0;
        __syncthreads();
        blk_walk<true>(S, g, t, parity, tt + g, t_stop, kBlkGroups, voff, idx16, idx32, blocked_tiles, val, tile_row, x, xs, 0u,
                       out, carry, head, vrows, pol_stream, pol_keep, use_tma != 0);
        tt = t_stop;
    }
    // ---- remainder tiles: interleaved share of [blocked_tiles, num_tiles), gathers through L1/L2 except for
    //      the hot prefix of x (only when there are no slabs: the slice buffer is free then) ----
    if (nbk == 0 && hot > 0) {
        for (u32 i = threadIdx.x; i < hot; i += kBlkThreads) xs[i] = ld_gather_f64(x + i, pol_keep);
        __syncthreads();
    }
    blk_walk<false>(S, g, t, parity, blocked_tiles + blockIdx.x * kBlkGroups + g, num_tiles, gridDim.x * kBlkGroups, voff,
                    idx16, idx32, blocked_tiles, val, tile_row, x, xs, nbk == 0 ? hot : 0u, out, carry, head, vrows,
                    pol_stream, pol_keep, use_tma != 0);
}

// y[r] = partial[0*R + r] + partial[1*R + r] + ... (slab order)
__global__ void k_slab_reduce(const double *__restrict__ partial, u64 rows, u32 nslabs, double *__restrict__ y) {
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; r < rows; r += stride) {
        double s = 0.0;
        for (u32 b = 0; b < nslabs; b++) s += partial[(u64)b * rows + r];
        y[r] = s;
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

int g_spmv_variant = -1;
int g_spmv_hot_entries = -1; // -1 auto
int g_spmv_tma = 1;          // 0: debugging path without bulk copies

namespace {
constexpr u32 kHotMaxEntries = 16384; // upper bound of the knob: 128 KB of the 227 KB
size_t hot_smem_bytes(u32 hot) { return sizeof(HotGroupSmem) * kHotGroups + (size_t)hot * sizeof(double); }
bool g_hot_attr_set = false;
} // namespace

namespace {
bool g_blk_attr_set = false;
size_t blk_smem_bytes(int groups, u32 xs_entries) { return sizeof(BlkGroupSmem) * groups + (size_t)xs_entries * sizeof(double); }

void spmv_blocked(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches) {
    const BlockedCsr &K = A.blk;
    if (!g_blk_attr_set) {
        LAS2_CUDA(cudaFuncSetAttribute(k_spmv_blocked<kBlkGroupsSlab>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)blk_smem_bytes(kBlkGroupsSlab, kBlockWidth)));
        LAS2_CUDA(cudaFuncSetAttribute(k_spmv_blocked<kBlkGroupsPlain>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)blk_smem_bytes(kBlkGroupsPlain, kHotPrefix)));
        g_blk_attr_set = true;
    }
    double *out = K.nslabs > 1 ? K.partial.get() : y;
    if (K.nbk > 0) {
        u32 grid = kNumSMs;
        if (grid * kBlkGroupsSlab > K.num_tiles) grid = cdiv(K.num_tiles, kBlkGroupsSlab);
        k_spmv_blocked<kBlkGroupsSlab><<<grid, kBlkGroupsSlab * kBlkLanes, blk_smem_bytes(kBlkGroupsSlab, kBlockWidth), s>>>(
            K.voff, K.idx16, K.idx32, K.val, K.tile_row, K.slab_tile_start, x, out, K.carry, K.head, (u32)K.vrows,
            (u32)K.cols, K.nbk, K.blocked_tiles, K.num_tiles, 0u, g_spmv_tma);
    } else {
        u32 hot = g_spmv_hot_entries >= 0 ? (u32)g_spmv_hot_entries : std::min<u32>(A.hot_entries, kPlainHotDefault);
        if (hot > kHotPrefix) hot = kHotPrefix;
        if (hot > K.cols) hot = (u32)K.cols;
        u32 grid = kNumSMs;
        if (grid * kBlkGroupsPlain > K.num_tiles) grid = cdiv(K.num_tiles, kBlkGroupsPlain);
        k_spmv_blocked<kBlkGroupsPlain><<<grid, kBlkGroupsPlain * kBlkLanes, blk_smem_bytes(kBlkGroupsPlain, hot), s>>>(
            K.voff, K.idx16, K.idx32, K.val, K.tile_row, K.slab_tile_start, x, out, K.carry, K.head, (u32)K.vrows,
            (u32)K.cols, K.nbk, K.blocked_tiles, K.num_tiles, hot, g_spmv_tma);
    }
    if (K.num_tiles > 1) {
        k_spmv_fixup<<<cdiv(K.num_tiles - 1, 128), 128, 0, s>>>(K.fix_row, K.fix_first, K.carry, K.head, out, K.num_tiles);
        if (launches) *launches += 1;
    }
    if (K.nslabs > 1) {
        unsigned rg = cdiv(K.rows, 256);
        if (rg > (unsigned)kNumSMs * 8) rg = kNumSMs * 8;
        k_slab_reduce<<<rg, 256, 0, s>>>(K.partial, K.rows, K.nslabs, y);
        if (launches) *launches += 1;
    }
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 1;
}
} // namespace

void spmv_chunk(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches); // chunk_spmv.cu

void spmv(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches) {
    if (A.rows == 0) return;
    if (A.chk.active && (g_spmv_variant < 0 || g_spmv_variant == 5)) {
        spmv_chunk(A, x, y, s, launches);
        return;
    }
    if (A.blk.active && (g_spmv_variant < 0 || g_spmv_variant == 4)) {
        spmv_blocked(A, x, y, s, launches);
        return;
    }
    // variant: -1 auto, 0/1/2 = one CTA per tile with 256/128/512 threads, 3 = persistent hot-prefix kernel
    int variant = g_spmv_variant;
    if (variant < 0) variant = (A.num_tiles >= 8u * kNumSMs) ? 3 : 0;
    switch (variant) {
    case 3: {
        u32 hot = g_spmv_hot_entries >= 0 ? (u32)g_spmv_hot_entries : A.hot_entries;
        if (hot > kHotMaxEntries) hot = kHotMaxEntries;
        if (hot > A.cols) hot = (u32)A.cols;
        if (!g_hot_attr_set) {
            LAS2_CUDA(cudaFuncSetAttribute(k_spmv_hot, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)hot_smem_bytes(kHotMaxEntries)));
            g_hot_attr_set = true;
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
