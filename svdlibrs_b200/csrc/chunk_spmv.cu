// chunk_spmv.cu -- the SpMV of the large-matrix path: y = A x over the "warp-chunk stream" (ChunkStream,
// device_csr.cuh).  Replaces the gather forms of SMat::svd_opa (reference src/lib.rs:2051-2055, 2087-2091); two
// launches per svd_opb (:835-836).
//
// What the earlier kernels taught (profiles/r1_spmv_iterations.md, r1_gather_microbench.md, gather_l1_sweep):
//   * HBM traffic was already minimal; the time went into ~57 thread instructions per non-zero (row-offset staging,
//     bitmap building with shared atomics, 3-way emission branches, four block barriers per 1024-nnz tile) and into
//     the L1TEX tag stage, which retires ~0.95 uncoalesced gathers per clock per SM -- but only while the
//     shared-memory carve-out stays <= 160 KB (0.43 per clock at 224 KB: misses in flight need L1 lines to land in).
// Hence this design:
//   * ONE WARP owns a chunk of 256 non-zeros end to end: no block barrier, no shared atomics, no row-offset loads.
//     Row starts travel inside the stream (bit 31 of the column index); the only per-chunk metadata is one u32.
//   * The chunk's 3072 bytes arrive by ONE bulk copy (cp.async.bulk + mbarrier) into the warp's own 2-deep ring,
//     off the LSU path, so the tag stage is left to the gathers.  16 warps x 2 x 3 KB + the hot prefix of x
//     (<= 32 KB) = 128 KB of shared memory, which keeps ~120 KB of L1 for gathers in flight.
//   * Each lane multiplies its 8 CONSECUTIVE elements and sums them serially into (piece before its first row
//     start, piece after); one warp-shuffle segmented scan stitches the lanes; rows that end inside the chunk are
//     written straight to y, the piece open at the chunk's start / end goes to head[] / carry[] and
//     k_chunk_fixup completes those rows in chunk order.  Every order is fixed by the data layout: deterministic.
//   * Chunks in which some lane holds two or more row starts (rows shorter than 8) take a branchy variant of the
//     per-lane step (marked at upload in bit 31 of chunk_row[]); everything else is shared.
//
// Algorithmic bytes per launch: 12*Z + 4*(R+1) + 8*C + 8*R.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "spmv.cuh"

namespace las2 {

void exclusive_scan_u32(u32 *data, u64 n, cudaStream_t s, UploadStats &st); // device_csr.cu

namespace {

constexpr int kChunkWarps = 16;
constexpr int kChunkThreads = kChunkWarps * 32;
constexpr u32 kChunkHotMax = 8192; // doubles of x kept in shared memory at most (64 KB)
constexpr u32 kChunkHotDefault = 4096;

// ------------------------------------------------------------------------------------------------ build

// ne[r] = 1 when row r has entries (ne has rows + 1 slots; the last is 0 so that the scan yields the total)
__global__ void k_chunk_nonempty(const u32 *__restrict__ off, u32 rows, u32 *__restrict__ ne) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; r <= rows; r += stride) ne[r] = (r < rows && off[r + 1] > off[r]) ? 1u : 0u;
}

// cstart[c] = first element of the c-th non-empty row, rowmap[c] = that row; cstart[nc] = nnz (the sentinel)
__global__ void k_chunk_compact(const u32 *__restrict__ off, const u32 *__restrict__ cid, u32 rows, u32 nnz,
                                u32 *__restrict__ cstart, u32 *__restrict__ rowmap) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; r <= rows; r += stride) {
        if (r == rows) { cstart[cid[rows]] = nnz; continue; }
        const u32 o = off[r];
        if (off[r + 1] > o) { cstart[cid[r]] = o; rowmap[cid[r]] = r; }
    }
}

__device__ __forceinline__ u64 chunk_val_slot(u64 e) { // position (in doubles) of element e's value in `data`
    const u64 c = e / kChunkNnz;
    const u32 j = (u32)(e % kChunkNnz), l = j / kChunkE, k = j % kChunkE;
    return c * (kChunkBytes / 8) + ((k >> 1) * 32 + l) * 2 + (k & 1);
}
__device__ __forceinline__ u64 chunk_idx_slot(u64 e) { // position (in u32) of element e's column index
    const u64 c = e / kChunkNnz;
    const u32 j = (u32)(e % kChunkNnz), l = j / kChunkE, k = j % kChunkE;
    return c * (kChunkBytes / 4) + (kChunkNnz * 2) + ((k >> 2) * 32 + l) * 4 + (k & 3);
}

__global__ void k_chunk_scatter(const u32 *__restrict__ idx, const double *__restrict__ val, u64 nnz, u64 padded,
                                uint4 *__restrict__ data) {
    double *dv = reinterpret_cast<double *>(data);
    u32 *di = reinterpret_cast<u32 *>(data);
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; e < padded; e += stride) {
        dv[chunk_val_slot(e)] = e < nnz ? val[e] : 0.0;
        di[chunk_idx_slot(e)] = e < nnz ? idx[e] : 0u;
    }
}

// chunk_row[b] = number of row starts before element b*256 (b = 0 .. num_chunks)
__global__ void k_chunk_rows(const u32 *__restrict__ cstart, u32 nc, u32 num_chunks, u32 *__restrict__ chunk_row) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > num_chunks) return;
    const u64 target = (u64)b * kChunkNnz;
    u32 lo = 0, hi = nc + 1; // lower_bound over cstart[0 .. nc]
    while (lo < hi) {
        const u32 mid = lo + ((hi - lo) >> 1);
        if ((u64)cstart[mid] < target) lo = mid + 1; else hi = mid;
    }
    chunk_row[b] = lo;
}

// row completed by the head of chunk b, and the chunk it started in (before the slow bits are set)
__global__ void k_chunk_fix_list(const u32 *__restrict__ cstart, const u32 *__restrict__ rowmap,
                                 const u32 *__restrict__ chunk_row, u32 num_chunks, u32 *__restrict__ fix_row,
                                 u32 *__restrict__ fix_first) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_chunks) return;
    u32 row = 0xffffffffu, first = 0;
    const u32 c0 = chunk_row[b];
    if (b > 0 && chunk_row[b + 1] > c0 && c0 > 0) {
        const u32 c = c0 - 1;               // the row that is open when chunk b begins
        first = cstart[c] / kChunkNnz;      // < b because cstart[c] < b*256
        row = rowmap ? rowmap[c] : c;
    }
    fix_row[b] = row;
    fix_first[b] = first;
}

// set the row-start bits (incl. the sentinel at element nnz) and mark chunks in which a lane sees >= 2 starts
__global__ void k_chunk_flags(const u32 *__restrict__ cstart, u32 nc, uint4 *__restrict__ data,
                              u32 *__restrict__ chunk_row) {
    u32 *di = reinterpret_cast<u32 *>(data);
    u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; c <= nc; c += stride) {
        const u32 e = cstart[c];
        di[chunk_idx_slot(e)] |= kChunkFlag; // one start per element: no two threads touch the same word
        if (c < nc && cstart[c + 1] / kChunkE == e / kChunkE) atomicOr(&chunk_row[e / kChunkNnz], kChunkFlag);
    }
}

// ------------------------------------------------------------------------------------------------ kernel

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, u32 bytes, unsigned long long *bar, u64 pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, u32 parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ double u2d(u32 lo, u32 hi) { return __hiloint2double((int)hi, (int)lo); }

constexpr size_t kChunkBarBytes = 256; // 16 warps x 2 mbarriers
size_t chunk_smem_bytes(u32 hot) { return kChunkBarBytes + (size_t)kChunkWarps * 2 * kChunkBytes + (size_t)hot * sizeof(double); }

template <bool HOT>
__global__ void __launch_bounds__(kChunkThreads, 1)
k_spmv_chunk(const uint4 *__restrict__ data, const u32 *__restrict__ chunk_row, const u32 *__restrict__ rowmap,
             const double *__restrict__ x, double *__restrict__ y, double *__restrict__ carry,
             double *__restrict__ head, u32 num_chunks, u32 hot) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned char *stages = smem_raw + kChunkBarBytes;
    const double *xs = reinterpret_cast<const double *>(stages + (size_t)kChunkWarps * 2 * kChunkBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();
    unsigned long long *bar = bars + warp * 2;
    unsigned char *stage = stages + (size_t)warp * 2 * kChunkBytes;

    const u32 total_warps = gridDim.x * kChunkWarps;
    u32 c = blockIdx.x * kChunkWarps + warp;
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (c < num_chunks) {
            mbar_expect_tx(&bar[0], kChunkBytes);
            bulk_load(stage, data + (u64)c * (kChunkBytes / 16), kChunkBytes, &bar[0], pol_stream);
        }
        if (c + total_warps < num_chunks) {
            mbar_expect_tx(&bar[1], kChunkBytes);
            bulk_load(stage + kChunkBytes, data + (u64)(c + total_warps) * (kChunkBytes / 16), kChunkBytes, &bar[1], pol_stream);
        }
    }
    if (HOT) {
        double *xw = const_cast<double *>(xs);
        for (u32 i = threadIdx.x; i < hot; i += kChunkThreads) xw[i] = ld_gather_f64(x + i, pol_keep);
        __syncthreads(); // the only block-wide barrier of the kernel
    } else {
        __syncwarp();
    }

    const u32 lt_mask = (1u << lane) - 1u;
    u32 phases = 0;
    // the chunk descriptor is fetched one iteration ahead: consecutive chunks of a warp are total_warps apart, so
    // every such load is a DRAM-latency miss (it was the top stall of the first version of this kernel)
    u32 cr_next = c < num_chunks ? chunk_row[c] : 0u;
    for (u32 it = 0; c < num_chunks; c += total_warps, it++) {
        const u32 sidx = it & 1u;
        const u32 cr = cr_next;
        if (c + total_warps < num_chunks) cr_next = chunk_row[c + total_warps];
        mbar_wait(&bar[sidx], (phases >> sidx) & 1u);
        phases ^= 1u << sidx;
        const uint4 *st = reinterpret_cast<const uint4 *>(stage + sidx * kChunkBytes);
        uint4 q[6];
#pragma unroll
        for (int j = 0; j < 6; j++) q[j] = st[j * 32 + lane];
        double v[kChunkE];
        u32 ci[kChunkE];
#pragma unroll
        for (int j = 0; j < 4; j++) { v[2 * j] = u2d(q[j].x, q[j].y); v[2 * j + 1] = u2d(q[j].z, q[j].w); }
        ci[0] = q[4].x; ci[1] = q[4].y; ci[2] = q[4].z; ci[3] = q[4].w;
        ci[4] = q[5].x; ci[5] = q[5].y; ci[6] = q[5].z; ci[7] = q[5].w;

        // ---- gathers: 8 independent loads in flight per lane ----
        double xv[kChunkE];
#pragma unroll
        for (int k = 0; k < (int)kChunkE; k++) {
            const u32 col = ci[k] & ~kChunkFlag;
            if (HOT) {
                if (col < hot) xv[k] = xs[col];
                else xv[k] = ld_gather_f64(x + col, pol_keep);
            } else {
                xv[k] = ld_gather_f64(x + col, pol_keep);
            }
        }

        const u32 c0 = cr & ~kChunkFlag;
        double hsum = 0.0, tsum = 0.0; // piece before the lane's first row start / piece after its last one
        bool seen = false;
        u32 base;                      // row starts in the lanes before this one
        if (!(cr & kChunkFlag)) {
            // ---- common case: at most one row start per lane ----
#pragma unroll
            for (int k = 0; k < (int)kChunkE; k++) {
                seen = seen || ((int)ci[k] < 0);
                if (seen) tsum = fma(v[k], xv[k], tsum);
                else hsum = fma(v[k], xv[k], hsum);
            }
            base = 0; // filled in below from the ballot
        } else {
            // ---- short rows: a lane may hold several starts; rows that begin and end inside the lane are written here ----
            u32 nf = 0;
#pragma unroll
            for (int k = 0; k < (int)kChunkE; k++) nf += ci[k] >> 31;
            u32 incl = nf;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            base = incl - nf;
            u32 j = 0;
#pragma unroll
            for (int k = 0; k < (int)kChunkE; k++) {
                if ((int)ci[k] < 0) {
                    if (seen) { // the row opened by this lane's previous start is complete
                        const u32 cr_ = c0 + base + j - 1;
                        y[rowmap ? rowmap[cr_] : cr_] = tsum;
                    }
                    seen = true;
                    tsum = 0.0;
                    j++;
                }
                if (seen) tsum = fma(v[k], xv[k], tsum);
                else hsum = fma(v[k], xv[k], hsum);
            }
            __syncwarp();
        }

        // ---- stitch the lanes: segmented inclusive scan of the pieces open at each lane's end ----
        const u32 mask = __ballot_sync(0xffffffffu, seen);
        if (!(cr & kChunkFlag)) base = __popc(mask & lt_mask);
        double xsum = seen ? tsum : hsum;
        const u32 le = mask & (lt_mask | (1u << lane));
        const int dist = le ? (lane - (31 - __clz(le))) : lane; // lanes back to the start of the open piece
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, xsum, o);
            if (dist >= o) xsum += up;
        }
        double cin = __shfl_up_sync(0xffffffffu, xsum, 1);
        if (lane == 0) cin = 0.0;

        // every lane's stage reads have long completed (their values fed the shuffles above): refill the slot
        if (lane == 0 && c + 2 * total_warps < num_chunks) {
            mbar_expect_tx(&bar[sidx], kChunkBytes);
            bulk_load(stage + sidx * kChunkBytes, data + (u64)(c + 2 * total_warps) * (kChunkBytes / 16), kChunkBytes,
                      &bar[sidx], pol_stream);
        }

        // ---- emission: a lane with a row start closes the piece that was open before it ----
        if (seen) {
            const double total = cin + hsum;
            if ((mask & lt_mask) == 0) {
                head[c] = total; // that piece began before this chunk
            } else {
                const u32 cr_ = c0 + base - 1;
                y[rowmap ? rowmap[cr_] : cr_] = total;
            }
        }
        if (lane == 31) carry[c] = xsum;
    }
}

// Rows that cross chunk boundaries: y[row] = carry[first] + ... + carry[b-1] + head[b].  One lane per chunk; spans of
// more than 32 chunks (rows with > 8192 entries) are summed by the whole warp: lane-strided partial sums in chunk
// order, then a fixed shuffle tree.
__global__ void __launch_bounds__(256)
k_chunk_fixup(const u32 *__restrict__ fix_row, const u32 *__restrict__ fix_first, const double *__restrict__ carry,
              const double *__restrict__ head, double *__restrict__ y, u32 num_chunks) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    u32 row = 0xffffffffu, first = 0;
    if (b < num_chunks) { row = fix_row[b]; first = fix_first[b]; }
    const bool mine = row != 0xffffffffu;
    const bool is_long = mine && (b - first > 32u);
    if (mine && !is_long) {
        double s = 0.0;
        for (u32 t = first; t < b; t++) s += carry[t];
        y[row] = s + head[b];
    }
    u32 todo = __ballot_sync(0xffffffffu, is_long);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const u32 bb = __shfl_sync(0xffffffffu, b, src);
        const u32 ff = __shfl_sync(0xffffffffu, first, src);
        double s = 0.0;
        for (u32 t = ff + lane; t < bb; t += 32) s += carry[t];
        s = warp_sum(s);
        if (lane == src) y[row] = s + head[b];
    }
}

int g_chunk_carve_hot = -2, g_chunk_carve_cold = -2; // carve-out currently set on the two instantiations

} // namespace

int g_chunk_smem_pad_kb = 0;  // tuning knob "spmv_smem_pad_kb": extra (unused) dynamic shared memory
int g_chunk_carveout_pct = -1; // tuning knob "spmv_carveout": shared-memory carve-out in percent of 228 KB (-1: exact fit)



// Builds X.chk from the natural CSR arrays of X (device-side, deterministic).
void build_chunk_stream(DeviceCsr &X, cudaStream_t s, UploadStats &st) {
    ChunkStream &K = X.chk;
    K = ChunkStream();
    // SVDLAS2_SPMV: unset = this format for matrices large enough to fill the persistent kernel; "chunk" = always;
    // anything else (e.g. "tile", "blocked") = never (the older kernels stay reachable for comparison).
    const char *want = std::getenv("SVDLAS2_SPMV");
    const bool force = want && std::string(want) == "chunk";
    if (want && want[0] && !force) return;
    if (X.nnz == 0 || X.rows == 0) return;
    if (!force && X.nnz < (u64)8 * kNumSMs * kSpmvTile) return;
    if (X.cols >= kChunkFlag || X.nnz + 2 * (u64)kChunkNnz >= ((u64)1 << 32)) return;

    const u32 rows = (u32)X.rows, nnz = (u32)X.nnz;
    const u32 num_chunks = nnz / kChunkNnz + 1; // always room for the sentinel row start at element nnz
    const u64 padded = (u64)num_chunks * kChunkNnz;
    auto grid_for = [](u64 n, unsigned per) { unsigned g = cdiv(n, per); return std::max(1u, std::min<unsigned>(g, kNumSMs * 32)); };

    // non-empty rows -> compact numbering
    DevBuf<u32> cid(rows + 1);
    k_chunk_nonempty<<<grid_for(rows + 1, 256), 256, 0, s>>>(X.off, rows, cid);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    exclusive_scan_u32(cid, (u64)rows + 1, s, st);
    u32 nc = 0;
    LAS2_CUDA(cudaMemcpyAsync(&nc, cid.get() + rows, sizeof(u32), cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    K.nonempty_rows = nc;
    K.has_empty_rows = nc != rows;
    DevBuf<u32> cstart_buf;
    const u32 *cstart = X.off.get(); // identity numbering: the row offsets are the starts, off[rows] = nnz the sentinel
    if (K.has_empty_rows) {
        cstart_buf.alloc((size_t)nc + 1);
        K.rowmap.alloc(std::max<u32>(nc, 1));
        k_chunk_compact<<<grid_for(rows + 1, 256), 256, 0, s>>>(X.off, cid, rows, nnz, cstart_buf, K.rowmap);
        LAS2_LAUNCH_CHECK();
        st.launches++;
        cstart = cstart_buf.get();
    }

    K.data.alloc(((size_t)num_chunks + 1) * (kChunkBytes / 16));
    k_chunk_scatter<<<grid_for(padded, 256 * 4), 256, 0, s>>>(X.idx, X.val, nnz, padded, K.data);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    K.chunk_row.alloc((size_t)num_chunks + 1);
    k_chunk_rows<<<cdiv((u64)num_chunks + 1, 256), 256, 0, s>>>(cstart, nc, num_chunks, K.chunk_row);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    K.fix_row.alloc(num_chunks);
    K.fix_first.alloc(num_chunks);
    k_chunk_fix_list<<<cdiv(num_chunks, 256), 256, 0, s>>>(cstart, K.has_empty_rows ? K.rowmap.get() : nullptr, K.chunk_row,
                                                          num_chunks, K.fix_row, K.fix_first);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    k_chunk_flags<<<grid_for((u64)nc + 1, 256), 256, 0, s>>>(cstart, nc, K.data, K.chunk_row);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    K.carry.alloc(num_chunks);
    K.head.alloc(num_chunks);
    LAS2_CUDA(cudaMemsetAsync(K.carry.get(), 0, num_chunks * sizeof(double), s));
    LAS2_CUDA(cudaMemsetAsync(K.head.get(), 0, num_chunks * sizeof(double), s));
    LAS2_CUDA(cudaStreamSynchronize(s)); // cid / cstart_buf go out of scope

    K.rows = X.rows; K.cols = X.cols; K.num_chunks = num_chunks;
    K.active = true;
}

void spmv_chunk(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches) {
    const ChunkStream &K = A.chk;
    u32 hot = g_spmv_hot_entries >= 0 ? (u32)g_spmv_hot_entries : std::min<u32>(A.hot_entries, kChunkHotDefault);
    hot = std::min<u32>(hot, kChunkHotMax);
    hot = (u32)std::min<u64>(hot, K.cols);
    const size_t smem = std::min<size_t>(chunk_smem_bytes(hot) + (size_t)g_chunk_smem_pad_kb * 1024, 227 * 1024);
    // The shared-memory carve-out is set explicitly to what this launch needs: left to itself the driver picks the
    // carve-out that would fit the most CTAs per SM (e.g. 196 KB for a 98 KB kernel), and the gathers then lose the
    // L1 lines their misses land in (scripts/micro/gather_l1_sweep.cu: 0.95 gathers/clk/SM with <= 132 KB carved
    // out, 0.43 with 228 KB).
    const int carve = g_chunk_carveout_pct >= 0 ? g_chunk_carveout_pct : (int)std::min<size_t>(100, (smem + 1024) * 100 / (228 * 1024) + 1);
    auto configure = [&](const void *fn, int &cached) {
        if (cached == carve) return;
        LAS2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        LAS2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        cached = carve;
    };
    if (K.has_empty_rows) LAS2_CUDA(cudaMemsetAsync(y, 0, K.rows * sizeof(double), s));
    const u32 grid = std::min<u32>(kNumSMs, cdiv(K.num_chunks, kChunkWarps));
    const u32 *rowmap = K.has_empty_rows ? K.rowmap.get() : nullptr;
    if (hot > 0) {
        configure((const void *)k_spmv_chunk<true>, g_chunk_carve_hot);
        k_spmv_chunk<true><<<grid, kChunkThreads, smem, s>>>(K.data, K.chunk_row, rowmap, x, y, K.carry, K.head, K.num_chunks, hot);
    } else {
        configure((const void *)k_spmv_chunk<false>, g_chunk_carve_cold);
        k_spmv_chunk<false><<<grid, kChunkThreads, smem, s>>>(K.data, K.chunk_row, rowmap, x, y, K.carry, K.head, K.num_chunks, 0u);
    }
    k_chunk_fixup<<<cdiv(K.num_chunks, 256), 256, 0, s>>>(K.fix_row, K.fix_first, K.carry, K.head, y, K.num_chunks);
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 2;
}

} // namespace las2
