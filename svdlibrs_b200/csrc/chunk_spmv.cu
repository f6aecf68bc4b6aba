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
//   * The chunk's 3072 bytes arrive by ONE bulk copy (cp.async.bulk + mbarrier) into the warp's own stage, off the
//     LSU path, so the tag stage is left to the gathers.  32 warps x 1 stage x 3 KB (96 KB; measured faster than
//     16 warps x 2 stages: the kernel is latency-bound and the other 31 warps hide a warp's copy) + the hot prefix
//     of x (<= 64 KB) stays <= 164 KB of shared memory, which keeps >= 90 KB of L1 for gathers in flight; the
//     carve-out is set explicitly per launch.
//   * Each lane multiplies its 8 CONSECUTIVE elements and sums them serially into (piece before its first row
//     start, piece after); one warp-shuffle segmented scan stitches the lanes; rows that end inside the chunk are
//     written straight to y, the piece open at the chunk's start / end goes to head[] / carry[] and
//     k_chunk_fixup completes those rows in chunk order.  Every order is fixed by the data layout: deterministic.
//   * Chunks in which some lane holds two or more row starts (rows shorter than 8) take a branchy variant of the
//     per-lane step (marked at upload in bit 31 of chunk_row[]); everything else is shared.
//   * Slab mode (MODE 2; chosen at upload for matrices of >= 20 M nnz whose gathered vector has no popular prefix and
//     whose (16384-column block, row) pieces hold >= 4 entries on average): the
//     stream is ordered (slab, row, column) with local column indices, each CTA takes a contiguous range of chunks
//     and keeps the slab's 128 KB slice of x in shared memory, so EVERY gather is an ld.shared (~0.33 wavefronts
//     instead of one L1 line; profiles/r1f_spmv_ncu_summary.md shows the L1TEX pipe, not HBM, bounding MODE 0/1).
//     The rows of the stream are then the non-empty (slab, row) pieces; partial sums are written by running piece
//     number (dense, coalesced) and k_slab_reduce adds the pieces of a row in slab order.
//
// Algorithmic bytes per launch: 12*Z + 4*(R+1) + 8*C + 8*R.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "spmv.cuh"
#include "vecops.cuh"

namespace las2 {

void exclusive_scan_u32(u32 *data, u64 n, cudaStream_t s, UploadStats &st); // device_csr.cu
void expand_row_ids(const u32 *off, u32 rows, u64 nnz, u32 *rowid, cudaStream_t s, UploadStats &st);

namespace {

// launch shapes: (warps per CTA, stages per warp).  Both fill 96 KB of shared memory with stages.  16 x 2 hides the
// bulk-copy latency inside each warp; 32 x 1 doubles the warps that hide each other's ALU / shuffle latencies (the
// slab mode is latency-bound: 41 % issue-active with 4 warps per scheduler, profiles/r1_spmv_iterations.md).
constexpr int kChunkStageBytesTotal = 32 * kChunkBytes;
constexpr u32 kChunkHotMax = 8192; // doubles of x kept in shared memory at most (64 KB)
constexpr u32 kChunkHotDefault = 8192; // cfg 2, B: 0.116 ms with 8192 entries, 0.122 ms with 4096

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

// ---- slab mode ----

// cnt[slab * rows + row] += 1 for every non-zero (integer atomics: the result does not depend on the order)
__global__ void k_piece_count(const u32 *__restrict__ idx, const u32 *__restrict__ rowid, u64 nnz, u64 rows, u32 width,
                              u32 *__restrict__ cnt) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) atomicAdd(&cnt[(u64)(idx[k] / width) * rows + rowid[k]], 1u);
}

// ne[v] = cnt[v] > 0 (ne has one more slot, set to 0, so that the scan leaves the total there)
__global__ void k_piece_nonempty(const u32 *__restrict__ cnt, u64 npieces, u32 *__restrict__ ne) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; v <= npieces; v += stride) ne[v] = (v < npieces && cnt[v] > 0) ? 1u : 0u;
}

// cstart[piece_id[v]] = padded stream position of piece v; cstart[total] = sentinel
__global__ void k_piece_starts(const u32 *__restrict__ pos, const u32 *__restrict__ piece_id, const u32 *__restrict__ shift,
                               u64 npieces, u64 rows, u32 sentinel, u32 *__restrict__ cstart) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; v <= npieces; v += stride) {
        if (v == npieces) { cstart[piece_id[npieces]] = sentinel; continue; }
        if (piece_id[v + 1] > piece_id[v]) cstart[piece_id[v]] = pos[v] + shift[v / rows];
    }
}

// ranked scatter into the (slab, row, column) order: inside a row the entries of one slab are contiguous (columns
// ascend), so the rank of entry k inside its piece is k minus the position of the piece's first entry in the row
__global__ void k_slab_scatter(const u32 *__restrict__ off, const u32 *__restrict__ idx, const double *__restrict__ val,
                               const u32 *__restrict__ rowid, u64 nnz, u64 rows, u32 width, int local,
                               const u32 *__restrict__ pos, const u32 *__restrict__ shift, uint4 *__restrict__ data) {
    double *dv = reinterpret_cast<double *>(data);
    u32 *di = reinterpret_cast<u32 *>(data);
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        const u32 c = idx[k], r = rowid[k];
        const u32 sl = c / width;
        const u32 first_col = sl * width;
        u32 lo = off[r], hi = (u32)k; // lower_bound of first_col in idx[off[r] .. k]
        while (lo < hi) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (idx[mid] < first_col) lo = mid + 1; else hi = mid;
        }
        const u64 e = (u64)pos[(u64)sl * rows + r] + shift[sl] + ((u32)k - lo);
        dv[chunk_val_slot(e)] = val[k];
        di[chunk_idx_slot(e)] = local ? c - first_col : c;
    }
}

// slot[v] = running number of piece v when it has entries, kNoPiece otherwise (one load per (slab, row) in the reduce)
constexpr u32 kNoPiece = 0xffffffffu;
__global__ void k_piece_slots(const u32 *__restrict__ piece_id, u64 npieces, u32 *__restrict__ slot) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; v < npieces; v += stride) {
        const u32 id = piece_id[v];
        slot[v] = piece_id[v + 1] > id ? id : kNoPiece;
    }
}

// y[r] = sum of the partials of the pieces (slab, r).  A CTA of 8 warps owns 32 consecutive rows; warp w sums slabs
// w, w+8, w+16, ... in ascending order with lane = row, so the slot loads of a warp are one 128-byte line and the
// partials it then fetches are (nearly) consecutive doubles; the eight per-warp sums of a row are added in warp order
// through shared memory.  Every order is fixed.  The slot numbers are static: with programmatic dependent launch
// they are fetched while the kernels before this one drain.
constexpr int kReduceWarps = 8;
constexpr int kReduceMaxPerWarp = 16; // slabs per warp held in registers; more slabs fall back to a loop
__global__ void __launch_bounds__(kReduceWarps * 32)
k_slab_reduce(const u32 *__restrict__ slot, const double *__restrict__ part, u64 rows, u32 nslabs, double *__restrict__ y) {
    cudaTriggerProgrammaticLaunchCompletion();
    __shared__ double sums[kReduceWarps][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 r = (u64)blockIdx.x * 32 + lane;
    const bool live = r < rows;
    double s = 0.0;
    if (nslabs <= kReduceWarps * kReduceMaxPerWarp) {
        u32 id[kReduceMaxPerWarp];
#pragma unroll
        for (int j = 0; j < kReduceMaxPerWarp; j++) {
            const u32 b = w + j * kReduceWarps;
            id[j] = (live && b < nslabs) ? slot[(u64)b * rows + r] : kNoPiece;
        }
        cudaGridDependencySynchronize();
        double p[kReduceMaxPerWarp];
#pragma unroll
        for (int j = 0; j < kReduceMaxPerWarp; j++) p[j] = id[j] != kNoPiece ? part[id[j]] : 0.0;
#pragma unroll
        for (int j = 0; j < kReduceMaxPerWarp; j++) s += p[j];
    } else {
        cudaGridDependencySynchronize();
        for (u32 b = w; live && b < nslabs; b += kReduceWarps) {
            const u32 id = slot[(u64)b * rows + r];
            if (id != kNoPiece) s += part[id];
        }
    }
    sums[w][lane] = s;
    __syncthreads();
    if (w == 0 && live) {
        double t = sums[0][lane];
#pragma unroll
        for (int k = 1; k < kReduceWarps; k++) t += sums[k][lane];
        y[r] = t;
    }
}

// The same for a FEW slabs (the wide-slab streams: 2..16 blocks of 2^20 columns): one thread per row adds the row's pieces in
// ascending slab order; the slot loads of a warp are one 128-byte line per slab and the partials it then fetches are nearly
// consecutive.  (k_slab_reduce spends eight warps and a shared-memory combine per 32 rows, which is what 610 slabs need and
// what made it instruction-bound at 10 slabs: 0.16 ms per SpMV at cfg 4, profiles/r2g_ncu_summary.md.)
constexpr int kFewSlabs = 16;
__global__ void __launch_bounds__(256)
k_slab_reduce_few(const u32 *__restrict__ slot, const double *__restrict__ part, u64 rows, u32 nslabs, double *__restrict__ y) {
    cudaTriggerProgrammaticLaunchCompletion();
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = r < rows;
    u32 id[kFewSlabs];
#pragma unroll
    for (int b = 0; b < kFewSlabs; b++) id[b] = (live && (u32)b < nslabs) ? slot[(u64)b * rows + r] : kNoPiece;
    cudaGridDependencySynchronize();
    double p[kFewSlabs];
#pragma unroll
    for (int b = 0; b < kFewSlabs; b++) p[b] = id[b] != kNoPiece ? part[id[b]] : 0.0;
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < kFewSlabs; b++) s += p[b];
    if (live) y[r] = s;
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

constexpr size_t kChunkBarBytes = 384; // up to 32 stage mbarriers + the slice mbarrier of the slab mode
size_t chunk_smem_bytes(u32 hot) { return kChunkBarBytes + (size_t)kChunkStageBytesTotal + (size_t)hot * sizeof(double); }

// One warp's walk over the chunks first, first + step, ... < stop.  MODE 0: every gather through L1; 1: columns below
// `hot` from the shared-memory prefix xs, the rest through L1; 2: every gather from the shared-memory slice xs.
template <int MODE, int STAGES>
__device__ __forceinline__ void chunk_walk(u32 first, u32 stop, u32 step, int lane, unsigned long long *bar,
                                           unsigned char *stage, u32 &phases, const double *xs, u32 hot,
                                           const uint4 *__restrict__ data, const u32 *__restrict__ chunk_row,
                                           const u32 *__restrict__ rowmap, const double *__restrict__ x,
                                           double *__restrict__ y, double *__restrict__ carry, double *__restrict__ head,
                                           u64 pol_stream, u64 pol_keep) {
    if (first >= stop) return;
    u32 c = first;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < STAGES; j++) {
            if (c + j * step < stop) {
                mbar_expect_tx(&bar[j], kChunkBytes);
                bulk_load(stage + j * kChunkBytes, data + (u64)(c + j * step) * (kChunkBytes / 16), kChunkBytes, &bar[j], pol_stream);
            }
        }
    }
    const u32 lt_mask = (1u << lane) - 1u;
    // the chunk descriptor is fetched one iteration ahead (consecutive chunks of a warp are far apart in memory)
    u32 cr_next = chunk_row[c];
    for (u32 it = 0; c < stop; c += step, it++) {
        const u32 sidx = STAGES == 1 ? 0u : (it & 1u);
        const u32 cr = cr_next;
        if (c + step < stop) cr_next = chunk_row[c + step];
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
            if (MODE == 2) {
                xv[k] = xs[col];
            } else if (MODE == 1) {
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
            // exclusive prefix of nf (0 .. 8) over the lanes, bit-sliced: four ballots instead of a five-step shuffle scan
            base = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) base += (u32)__popc(__ballot_sync(0xffffffffu, (nf >> b) & 1u) & lt_mask) << b;
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
        // Every lane's reads of the stage have completed: the multiply-adds above consumed all six loaded vectors and
        // the ballot is issued for the whole warp only once they have.  Refill the slot for the chunk STAGES ahead.
        if (lane == 0 && c + STAGES * step < stop) {
            mbar_expect_tx(&bar[sidx], kChunkBytes);
            bulk_load(stage + sidx * kChunkBytes, data + (u64)(c + STAGES * step) * (kChunkBytes / 16), kChunkBytes, &bar[sidx],
                      pol_stream);
        }
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

template <int MODE, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmv_chunk(const uint4 *__restrict__ data, const u32 *__restrict__ chunk_row, const u32 *__restrict__ rowmap,
             const double *__restrict__ x, double *__restrict__ y, double *__restrict__ carry,
             double *__restrict__ head, u32 num_chunks, u32 hot, const u32 *__restrict__ slab_chunk_start, u32 nslabs,
             u32 cols) {
    cudaTriggerProgrammaticLaunchCompletion(); // the fix-up kernel may take over SMs as CTAs of this grid retire
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned char *stages = smem_raw + kChunkBarBytes;
    static_assert(WARPS * STAGES * kChunkBytes == kChunkStageBytesTotal, "stages fill 96 KB");
    constexpr int kChunkWarps = WARPS, kChunkThreads = WARPS * 32;
    double *xs = reinterpret_cast<double *>(stages + (size_t)kChunkStageBytesTotal);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();
    unsigned long long *bar = bars + warp * STAGES;
    unsigned char *stage = stages + (size_t)warp * STAGES * kChunkBytes;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < STAGES; j++) mbar_init(&bar[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    u32 phases = 0;
    unsigned long long *slice_bar = bars + WARPS * STAGES;
    u32 slice_phase = 0;
    if (MODE == 2) {
        if (threadIdx.x == 0) {
            mbar_init(slice_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    // Programmatic dependent launch: up to here only this kernel's own shared memory was touched; x (written by the
    // kernel before this one) is read and y / carry / head are written only after the preceding grid has completed.
    cudaGridDependencySynchronize();

    if (MODE == 2) {
        // contiguous share of the chunks: one or two slabs per CTA, i.e. one or two loads of a 128 KB slice of x
        const u32 lo = (u32)(((u64)num_chunks * blockIdx.x) / gridDim.x);
        const u32 hi = (u32)(((u64)num_chunks * (blockIdx.x + 1)) / gridDim.x);
        u32 sl = 0;
        for (u32 cur = lo; cur < hi;) {
            while (sl + 1 < nslabs && slab_chunk_start[sl + 1] <= cur) sl++;
            const u32 stop = min(hi, slab_chunk_start[sl + 1]);
            __syncthreads(); // every warp is done with the previous slice
            const u32 col0 = sl * kSlabWidth;
            const u32 n = min(kSlabWidth, cols - col0);
            if (threadIdx.x == 0) {
                // one bulk copy for the slice (an even number of doubles: 16-byte granularity), the odd one by hand
                const u32 bytes = (n & ~1u) * (u32)sizeof(double);
                if (bytes) {
                    mbar_expect_tx(slice_bar, bytes);
                    bulk_load(xs, x + col0, bytes, slice_bar, pol_keep);
                }
                if (n & 1u) xs[n - 1] = ld_gather_f64(x + col0 + n - 1, pol_keep);
            }
            if (n > 1) { mbar_wait(slice_bar, slice_phase); slice_phase ^= 1u; }
            __syncthreads();
            chunk_walk<2, STAGES>(cur + warp, stop, kChunkWarps, lane, bar, stage, phases, xs, 0u, data, chunk_row, rowmap, x, y,
                          carry, head, pol_stream, pol_keep);
            cur = stop;
        }
    } else {
        if (MODE == 1) {
            for (u32 i = threadIdx.x; i < hot; i += kChunkThreads) xs[i] = ld_gather_f64(x + i, pol_keep);
            __syncthreads(); // the only block-wide barrier of this mode
        }
        chunk_walk<MODE, STAGES>(blockIdx.x * kChunkWarps + warp, num_chunks, gridDim.x * kChunkWarps, lane, bar, stage, phases, xs,
                         hot, data, chunk_row, rowmap, x, y, carry, head, pol_stream, pol_keep);
    }
}

// Rows that cross chunk boundaries: y[row] = carry[first] + ... + carry[b-1] + head[b].  One lane per chunk; the common
// span of one chunk needs no loop (all loads issued together); spans of more than 32 chunks (rows with > 8192 entries)
// are summed by the whole warp: lane-strided partial sums in chunk order, then a fixed shuffle tree.  Launched with
// programmatic stream serialisation: the (static) fix-up lists are fetched while the SpMV kernel drains.
__global__ void __launch_bounds__(256)
k_chunk_fixup(const u32 *__restrict__ fix_row, const u32 *__restrict__ fix_first, const double *__restrict__ carry,
              const double *__restrict__ head, double *__restrict__ y, u32 num_chunks) {
    cudaTriggerProgrammaticLaunchCompletion(); // dependents may be scheduled as soon as SM resources free up
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    u32 row = 0xffffffffu, first = 0;
    if (b < num_chunks) { row = fix_row[b]; first = fix_first[b]; }
    cudaGridDependencySynchronize();
    const bool mine = row != 0xffffffffu;
    const bool is_long = mine && (b - first > 32u);
    if (mine && !is_long) {
        const double h = head[b];
        double s = carry[first];
        for (u32 t = first + 1; t < b; t++) s += carry[t];
        y[row] = s + h;
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

// ------------------------------------------------------------------------------------------------ 4 right-hand sides
// Y4 = A X4 for FOUR vectors at once (the sigma / U tail of ritvec, reference src/lib.rs:1839-1859, applies the operator to
// every accepted Ritz vector: d SpMVs that re-read the matrix d times).  X4 is row-interleaved (cols x 4 doubles), so the
// gather of one non-zero is ONE 32-byte sector that serves four vectors -- the L1TEX tag stage, which bounds the SpMV at
// ~1 gather per clock per SM, is paid once per four results -- and the 12 B/nnz stream is read once per four vectors.
// Same chunk stream, same per-lane order of the multiply-adds and the same segmented scan as k_spmv_chunk, applied to each
// of the four columns: every column of Y4 is bit-identical to the SpMV of that column.  16 warps x 2 stages (the 4-wide
// accumulators and gathers need ~120 registers per thread).
struct D4 { double v[4]; };
__device__ __forceinline__ D4 ld_gather4(const double *p, u64 pol) {
    D4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_row4(double *p, const D4 &a) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.v[0]), "d"(a.v[1]), "d"(a.v[2]), "d"(a.v[3]) : "memory");
}

template <int STAGES>
__device__ __forceinline__ void chunk_walk4(u32 first, u32 stop, u32 step, int lane, unsigned long long *bar,
                                            unsigned char *stage, u32 &phases, const uint4 *__restrict__ data,
                                            const u32 *__restrict__ chunk_row, const u32 *__restrict__ rowmap,
                                            const double *__restrict__ x4, double *__restrict__ y4,
                                            double *__restrict__ carry4, double *__restrict__ head4, u64 pol_stream, u64 pol_keep) {
    if (first >= stop) return;
    u32 c = first;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < STAGES; j++) {
            if (c + j * step < stop) {
                mbar_expect_tx(&bar[j], kChunkBytes);
                bulk_load(stage + j * kChunkBytes, data + (u64)(c + j * step) * (kChunkBytes / 16), kChunkBytes, &bar[j], pol_stream);
            }
        }
    }
    const u32 lt_mask = (1u << lane) - 1u;
    u32 cr_next = chunk_row[c];
    for (u32 it = 0; c < stop; c += step, it++) {
        const u32 sidx = STAGES == 1 ? 0u : (it & 1u);
        const u32 cr = cr_next;
        if (c + step < stop) cr_next = chunk_row[c + step];
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

        const u32 c0 = cr & ~kChunkFlag;
        D4 hsum = {{0.0, 0.0, 0.0, 0.0}}, tsum = {{0.0, 0.0, 0.0, 0.0}};
        bool seen = false;
        u32 base = 0;
        const bool slow = (cr & kChunkFlag) != 0;
        if (slow) {
            u32 nf = 0;
#pragma unroll
            for (int k = 0; k < (int)kChunkE; k++) nf += ci[k] >> 31;
#pragma unroll
            for (int b = 0; b < 4; b++) base += (u32)__popc(__ballot_sync(0xffffffffu, (nf >> b) & 1u) & lt_mask) << b;
        }
        u32 j = 0;
        // two halves of four elements: 4 x 32-byte gathers in flight per lane
#pragma unroll
        for (int h = 0; h < 2; h++) {
            D4 xv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) xv[k] = ld_gather4(x4 + (u64)(ci[4 * h + k] & ~kChunkFlag) * 4, pol_keep);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int e = 4 * h + k;
                if ((int)ci[e] < 0) {
                    if (slow) {
                        if (seen) { // the row opened by this lane's previous start is complete
                            const u32 cr_ = c0 + base + j - 1;
                            st_row4(y4 + (u64)(rowmap ? rowmap[cr_] : cr_) * 4, tsum);
                        }
#pragma unroll
                        for (int r = 0; r < 4; r++) tsum.v[r] = 0.0;
                        j++;
                    }
                    seen = true;
                }
                if (seen) {
#pragma unroll
                    for (int r = 0; r < 4; r++) tsum.v[r] = fma(v[e], xv[k].v[r], tsum.v[r]);
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) hsum.v[r] = fma(v[e], xv[k].v[r], hsum.v[r]);
                }
            }
        }
        if (slow) __syncwarp();

        const u32 mask = __ballot_sync(0xffffffffu, seen);
        if (lane == 0 && c + STAGES * step < stop) {
            mbar_expect_tx(&bar[sidx], kChunkBytes);
            bulk_load(stage + sidx * kChunkBytes, data + (u64)(c + STAGES * step) * (kChunkBytes / 16), kChunkBytes, &bar[sidx],
                      pol_stream);
        }
        if (!slow) base = __popc(mask & lt_mask);
        D4 xsum;
#pragma unroll
        for (int r = 0; r < 4; r++) xsum.v[r] = seen ? tsum.v[r] : hsum.v[r];
        const u32 le = mask & (lt_mask | (1u << lane));
        const int dist = le ? (lane - (31 - __clz(le))) : lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double up = __shfl_up_sync(0xffffffffu, xsum.v[r], o);
                if (dist >= o) xsum.v[r] += up;
            }
        }
        D4 cin;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            cin.v[r] = __shfl_up_sync(0xffffffffu, xsum.v[r], 1);
            if (lane == 0) cin.v[r] = 0.0;
        }
        if (seen) {
            D4 total;
#pragma unroll
            for (int r = 0; r < 4; r++) total.v[r] = cin.v[r] + hsum.v[r];
            if ((mask & lt_mask) == 0) {
                st_row4(head4 + (u64)c * 4, total);
            } else {
                const u32 cr_ = c0 + base - 1;
                st_row4(y4 + (u64)(rowmap ? rowmap[cr_] : cr_) * 4, total);
            }
        }
        if (lane == 31) st_row4(carry4 + (u64)c * 4, xsum);
    }
}

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_spmm4_chunk(const uint4 *__restrict__ data, const u32 *__restrict__ chunk_row, const u32 *__restrict__ rowmap,
              const double *__restrict__ x4, double *__restrict__ y4, double *__restrict__ carry4,
              double *__restrict__ head4, u32 num_chunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned char *stages = smem_raw + kChunkBarBytes;
    static_assert(WARPS * STAGES * kChunkBytes == kChunkStageBytesTotal, "stages fill 96 KB");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 pol_stream = l2_policy_evict_first();
    const u64 pol_keep = l2_policy_evict_last();
    unsigned long long *bar = bars + warp * STAGES;
    unsigned char *stage = stages + (size_t)warp * STAGES * kChunkBytes;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < STAGES; j++) mbar_init(&bar[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    u32 phases = 0;
    chunk_walk4<STAGES>(blockIdx.x * WARPS + warp, num_chunks, gridDim.x * WARPS, lane, bar, stage, phases, data, chunk_row,
                        rowmap, x4, y4, carry4, head4, pol_stream, pol_keep);
}

// rows that cross chunk boundaries, four columns at a time (same order of the adds as k_chunk_fixup)
__global__ void __launch_bounds__(256)
k_chunk_fixup4(const u32 *__restrict__ fix_row, const u32 *__restrict__ fix_first, const double *__restrict__ carry4,
               const double *__restrict__ head4, double *__restrict__ y4, u32 num_chunks) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    u32 row = 0xffffffffu, first = 0;
    if (b < num_chunks) { row = fix_row[b]; first = fix_first[b]; }
    const bool mine = row != 0xffffffffu;
    const bool is_long = mine && (b - first > 32u);
    if (mine && !is_long) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double s = carry4[(u64)first * 4 + r];
            for (u32 t = first + 1; t < b; t++) s += carry4[(u64)t * 4 + r];
            y4[(u64)row * 4 + r] = s + head4[(u64)b * 4 + r];
        }
    }
    u32 todo = __ballot_sync(0xffffffffu, is_long);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const u32 bb = __shfl_sync(0xffffffffu, b, src);
        const u32 ff = __shfl_sync(0xffffffffu, first, src);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double s = 0.0;
            for (u32 t = ff + lane; t < bb; t += 32) s += carry4[(u64)t * 4 + r];
            s = warp_sum(s);
            if (lane == src) y4[(u64)row * 4 + r] = s + head4[(u64)b * 4 + r];
        }
    }
}

// x4[i*4 + r] = V[(r) * ld + i] for r < nv (zero for the missing columns of the last batch)
__global__ void k_pack4(const double *__restrict__ V, u64 ld, u64 n, u32 nv, double *__restrict__ x4) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        D4 o;
#pragma unroll
        for (int r = 0; r < 4; r++) o.v[r] = (u32)r < nv ? V[(u64)r * ld + i] : 0.0;
        st_row4(x4 + i * 4, o);
    }
}

// U[r * ldu + i] = y4[i*4 + r] and per-CTA partials of sum U[r]^2 (parts[r * kMaxParts + cta]), r < nv
__global__ void __launch_bounds__(kVecThreads)
k_unpack4_sumsq(const double *__restrict__ y4, u64 m, u32 nv, double *__restrict__ U, u64 ldu, double *__restrict__ parts,
                u32 parts_stride) {
    __shared__ double sm[4][kVecThreads / 32];
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (; i < m; i += stride) {
        const double2 a = *reinterpret_cast<const double2 *>(y4 + i * 4), b = *reinterpret_cast<const double2 *>(y4 + i * 4 + 2);
        const double val[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if ((u32)r < nv) {
                U[(u64)r * ldu + i] = val[r];
                acc[r] = fma(val[r], val[r], acc[r]);
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const double w = warp_sum(acc[r]);
        if (lane == 0) sm[r][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kVecThreads / 32; w++) t += sm[threadIdx.x][w];
        parts[(u64)threadIdx.x * parts_stride + blockIdx.x] = t;
    }
}

int g_chunk_carve[kMaxDevicesTracked][6]; // carve-out (+1) currently set on each instantiation, per device (0 = not yet)

} // namespace



namespace {

// Programmatic dependent launch of the SpMV kernel chain (~2 % per opb at cfg 2; Nsight Compute lists and times the
// kernels either way).  Knob "spmv_pdl" or SVDLAS2_PDL=0 turns it off.
bool chunk_pdl_enabled(int knob) {
    static const int env = [] { const char *e = std::getenv("SVDLAS2_PDL"); return e ? std::atoi(e) : -1; }();
    return knob >= 0 ? knob != 0 : env != 0;
}

// Launch with programmatic stream serialisation: the kernel may start while its predecessor in the stream drains; it
// must call cudaGridDependencySynchronize() before touching anything the predecessor wrote (or still reads).
template <typename... KArgs, typename... Args>
void launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    if (!pdl) { // the plain launch path: visible to every profiler
        kernel<<<grid, block, smem, s>>>(static_cast<KArgs>(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    LAS2_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}

unsigned grid_for(u64 n, unsigned per) {
    unsigned g = cdiv(n, per);
    return std::max(1u, std::min<unsigned>(g, kNumSMs * 32));
}

// chunk table, fix-up lists, row-start bits and carry buffers for a stream whose rows start at cstart[0 .. nc-1]
// (ascending positions in the padded stream; cstart[nc] = the sentinel right behind the last real element)
void finalize_chunks(ChunkStream &K, const u32 *cstart, const u32 *rowmap, u32 nc, u32 num_chunks, cudaStream_t s,
                     UploadStats &st) {
    K.chunk_row.alloc((size_t)num_chunks + 1);
    k_chunk_rows<<<cdiv((u64)num_chunks + 1, 256), 256, 0, s>>>(cstart, nc, num_chunks, K.chunk_row);
    LAS2_LAUNCH_CHECK();
    K.fix_row.alloc(num_chunks);
    K.fix_first.alloc(num_chunks);
    k_chunk_fix_list<<<cdiv(num_chunks, 256), 256, 0, s>>>(cstart, rowmap, K.chunk_row, num_chunks, K.fix_row, K.fix_first);
    LAS2_LAUNCH_CHECK();
    k_chunk_flags<<<grid_for((u64)nc + 1, 256), 256, 0, s>>>(cstart, nc, K.data, K.chunk_row);
    LAS2_LAUNCH_CHECK();
    st.launches += 3;
    K.carry.alloc(num_chunks);
    K.head.alloc(num_chunks);
    LAS2_CUDA(cudaMemsetAsync(K.carry.get(), 0, num_chunks * sizeof(double), s));
    LAS2_CUDA(cudaMemsetAsync(K.head.get(), 0, num_chunks * sizeof(double), s));
    K.num_chunks = num_chunks;
}

// Slab mode.  Returns false (leaving K untouched apart from scratch) when the matrix is too sparse per piece.
// A CSR matrix (or a row range of one) as the builders see it: offsets rebased to 0, arrays borrowed.
struct CsrView {
    u64 rows, cols, nnz;
    const u32 *off, *idx;
    const double *val;
};

// Slab-ordered stream.  local: slab-local column indices, width kSlabWidth (all gathers from shared memory); otherwise global
// indices and any width (gathers through L1, the order only blocks the gathered vector for L2).  Returns false (leaving K
// untouched apart from scratch) when a (slab, row) piece holds fewer than `min_per_piece` entries on average.
bool build_slab_stream(const CsrView &X, ChunkStream &K, u32 width, bool local, bool force, double min_per_piece, cudaStream_t s,
                       UploadStats &st) {
    const u64 rows = X.rows, nnz = X.nnz;
    const u32 nslabs = cdiv(X.cols, width);
    const u64 npieces = (u64)nslabs * rows;
    if (nslabs < 2 || npieces + 1 >= ((u64)1 << 28)) return false;

    DevBuf<u32> rowid(nnz);
    expand_row_ids(X.off, (u32)rows, nnz, rowid, s, st);
    DevBuf<u32> pos(npieces + 1);
    LAS2_CUDA(cudaMemsetAsync(pos.get(), 0, (npieces + 1) * sizeof(u32), s));
    k_piece_count<<<grid_for(nnz, 256 * 4), 256, 0, s>>>(X.idx, rowid, nnz, rows, width, pos);
    LAS2_LAUNCH_CHECK();
    K.piece_id.alloc(npieces + 1);
    k_piece_nonempty<<<grid_for(npieces + 1, 256 * 4), 256, 0, s>>>(pos, npieces, K.piece_id);
    LAS2_LAUNCH_CHECK();
    st.launches += 2;
    exclusive_scan_u32(K.piece_id, npieces + 1, s, st);
    u32 nc = 0;
    LAS2_CUDA(cudaMemcpyAsync(&nc, K.piece_id.get() + npieces, sizeof(u32), cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    if (nc == 0 || (!force && (double)nnz < min_per_piece * (double)nc)) { K.piece_id.release(); return false; }

    // positions: entries per piece -> exclusive scan; every slab then starts on a chunk boundary
    exclusive_scan_u32(pos, npieces + 1, s, st);
    std::vector<u32> starts(nslabs + 1);
    LAS2_CUDA(cudaMemcpy2DAsync(starts.data(), sizeof(u32), pos.get(), rows * sizeof(u32), sizeof(u32), nslabs + 1,
                                cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    std::vector<u32> shift(nslabs), chunk_start(nslabs + 1);
    u64 p = 0, sentinel = 0;
    for (u32 i = 0; i < nslabs; i++) {
        shift[i] = (u32)(p - starts[i]);
        chunk_start[i] = (u32)(p / kChunkNnz);
        sentinel = p + (starts[i + 1] - starts[i]);
        p = (sentinel + kChunkNnz - 1) / kChunkNnz * kChunkNnz;
    }
    if (sentinel + 2 * (u64)kChunkNnz >= ((u64)1 << 32)) { K.piece_id.release(); return false; }
    const u32 num_chunks = (u32)(sentinel / kChunkNnz) + 1; // the sentinel's chunk belongs to the last slab
    chunk_start[nslabs] = num_chunks;
    DevBuf<u32> dshift(nslabs);
    LAS2_CUDA(cudaMemcpyAsync(dshift.get(), shift.data(), nslabs * sizeof(u32), cudaMemcpyHostToDevice, s));
    K.slab_chunk_start.alloc(nslabs + 1);
    LAS2_CUDA(cudaMemcpyAsync(K.slab_chunk_start.get(), chunk_start.data(), (nslabs + 1) * sizeof(u32), cudaMemcpyHostToDevice, s));

    DevBuf<u32> cstart((size_t)nc + 1);
    k_piece_starts<<<grid_for(npieces + 1, 256 * 4), 256, 0, s>>>(pos, K.piece_id, dshift, npieces, rows, (u32)sentinel, cstart);
    LAS2_LAUNCH_CHECK();
    K.data.alloc(((size_t)num_chunks + 1) * (kChunkBytes / 16));
    LAS2_CUDA(cudaMemsetAsync(K.data.get(), 0, ((size_t)num_chunks + 1) * kChunkBytes, s)); // padding: value 0, column 0
    k_slab_scatter<<<grid_for(nnz, 256 * 2), 256, 0, s>>>(X.off, X.idx, X.val, rowid, nnz, rows, width, local ? 1 : 0, pos, dshift, K.data);
    LAS2_LAUNCH_CHECK();
    st.launches += 2;
    finalize_chunks(K, cstart, nullptr, nc, num_chunks, s, st);
    {
        DevBuf<u32> slots(npieces);
        k_piece_slots<<<grid_for(npieces, 256 * 4), 256, 0, s>>>(K.piece_id, npieces, slots);
        LAS2_LAUNCH_CHECK();
        st.launches++;
        LAS2_CUDA(cudaStreamSynchronize(s));
        K.piece_id = std::move(slots); // from here on: slot numbers (kNoPiece for empty pieces)
    }
    K.part.alloc(nc);
    LAS2_CUDA(cudaMemsetAsync(K.part.get(), 0, (size_t)nc * sizeof(double), s));
    LAS2_CUDA(cudaStreamSynchronize(s)); // scratch and host vectors go out of scope
    K.slabs = true;
    K.slab_local = local;
    K.slab_width = width;
    K.nslabs = nslabs;
    K.nonempty_rows = nc;
    K.stream_bytes = (u64)num_chunks * kChunkBytes + 4 * (u64)num_chunks + 16 * (u64)nc + 4 * (npieces + 1) + 8 * X.cols * 2 + 8 * rows;
    return true;
}

void build_plain_stream(const CsrView &X, ChunkStream &K, cudaStream_t s, UploadStats &st) {
    const u32 rows = (u32)X.rows, nnz = (u32)X.nnz;
    const u32 num_chunks = nnz / kChunkNnz + 1; // always room for the sentinel row start at element nnz
    const u64 padded = (u64)num_chunks * kChunkNnz;

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
    const u32 *cstart = X.off; // identity numbering: the row offsets are the starts, off[rows] = nnz the sentinel
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
    finalize_chunks(K, cstart, K.has_empty_rows ? K.rowmap.get() : nullptr, nc, num_chunks, s, st);
    LAS2_CUDA(cudaStreamSynchronize(s)); // cid / cstart_buf go out of scope
    K.stream_bytes = (u64)num_chunks * kChunkBytes + 4 * (u64)num_chunks + 8 * X.cols + 8 * X.rows;
}

__global__ void k_rebase_offsets(const u32 *__restrict__ off, u32 r0, u32 n, u32 *__restrict__ out) {
    const u32 base = off[r0];
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; i <= n; i += stride) out[i] = off[r0 + i] - base;
}

constexpr u64 kWideSlabWidth = (u64)1 << 20;        // 8 MB of the gathered vector per wide slab
constexpr u64 kWideSlabMinCols = (u64)3 << 19;      // gathered vectors beyond 12 MB are blocked
constexpr u64 kSlabMinNnz = 20000000;               // per-launch cost of the slab modes (slice loads, reduce kernel)

int env_int(const char *name, int dflt) { const char *e = std::getenv(name); return (e && e[0]) ? std::atoi(e) : dflt; }

// the best stream for one row range: slab-local when forced / dense enough per piece, else wide slabs for a huge gathered
// vector, else the plain stream (with the matrix's hot prefix, if it has one)
void build_range(const CsrView &V, ChunkStream &K, u32 hot_entries, bool try_local, bool force_local, cudaStream_t s, UploadStats &st) {
    K = ChunkStream();
    bool done = false;
    const int wide_mode = env_int("SVDLAS2_WIDE_SLABS", -1); // -1 auto, 0 never, 1 whenever there are >= 2 slabs
    if (try_local || force_local) done = build_slab_stream(V, K, kSlabWidth, true, force_local, 4.0, s, st);
    if (!done && wide_mode != 0 && hot_entries == 0 && V.cols >= 2 * kWideSlabWidth &&
        (wide_mode == 1 || (V.cols >= kWideSlabMinCols && V.nnz >= kSlabMinNnz))) {
        K = ChunkStream();
        done = build_slab_stream(V, K, (u32)kWideSlabWidth, false, wide_mode == 1, 4.0, s, st);
    }
    if (!done) { K = ChunkStream(); build_plain_stream(V, K, s, st); K.hot_entries = hot_entries; }
    K.rows = V.rows; K.cols = V.cols;
    K.active = true;
}

} // namespace

// Builds X.chk (and X.chk2) from the natural CSR arrays of X (device-side, deterministic).
void build_chunk_stream(DeviceCsr &X, cudaStream_t s, UploadStats &st) {
    X.chk = ChunkStream();
    X.chk2 = ChunkStream();
    // SVDLAS2_SPMV: unset = this format for matrices large enough to fill the persistent kernel; "chunk" = always;
    // anything else (e.g. "tile") = never (the tile kernels stay reachable for comparison).
    // SVDLAS2_SLABS: unset = slab mode where a (slab, row) piece holds >= 4 entries on average; 0 = never; 1 = always.
    // SVDLAS2_SPLIT: unset / 1 = a frequency-ordered B' is cut into heavy rows (slab mode) and light rows; 0 = never.
    const char *want = std::getenv("SVDLAS2_SPMV");
    const bool force = want && std::string(want) == "chunk";
    if (want && want[0] && !force) return;
    if (X.nnz == 0 || X.rows == 0) return;
    if (!force && X.nnz < (u64)8 * kNumSMs * kSpmvTile) return;
    if (X.cols >= kChunkFlag || X.rows >= kChunkFlag || X.nnz + 2 * (u64)kChunkNnz >= ((u64)1 << 32)) return;
    const char *sl = std::getenv("SVDLAS2_SLABS");
    const char slab_mode = (sl && sl[0]) ? sl[0] : 'a';
    const CsrView whole{X.rows, X.cols, X.nnz, X.off.get(), X.idx.get(), X.val.get()};
    // Slabs only when the hot prefix cannot do the job (cfg 2: B, 65 % of whose gathers hit the first 8192 entries of x,
    // runs 0.12 ms with the prefix and 0.19 ms in slab mode; B', whose gathered vector t has no popular prefix, 0.19 ms
    // without slabs and 0.15 ms with them), and only for matrices (shards) big enough to amortise the mode's fixed cost.
    const bool slab_candidate = slab_mode == '1' || (slab_mode != '0' && X.hot_entries == 0 && X.nnz >= kSlabMinNnz);

    // ---- split of a frequency-ordered B': rows whose (slab, row) pieces are dense enough run in slab-local mode, the light
    // rows apart (wide slabs for a huge gathered vector, else plain): at cfg 4 the 35 k heaviest rows of B' hold 56 % of
    // the non-zeros with >= 6 entries per piece, while the whole matrix averages < 1 ----
    if (X.rows_sorted_desc && slab_candidate && slab_mode != '1' && env_int("SVDLAS2_SPLIT", 1) != 0 && X.rows >= 4096) {
        const u32 nslabs16 = cdiv(X.cols, kSlabWidth);
        std::vector<u32> off(X.rows + 1);
        LAS2_CUDA(cudaMemcpyAsync(off.data(), X.off.get(), (X.rows + 1) * sizeof(u32), cudaMemcpyDeviceToHost, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
        // rows are ordered by GLOBAL popularity; a shard's own lengths follow it only on average, hence windows of 1024 rows
        const double thr = (double)env_int("SVDLAS2_SPLIT_PER_PIECE", 4) * nslabs16; // cfg 4: 3 / 4 / 6 / 8 / 12 -> B' 2.99 / 3.00 / 3.03 / 3.05 / 3.11 ms
        u64 H = 0;
        constexpr u64 win = 1024;
        while (H + win <= X.rows && (double)(off[H + win] - off[H]) >= thr * win) H += win;
        const u64 nnz_heavy = off[H], nnz_light = X.nnz - nnz_heavy;
        if (H > 0 && H < X.rows && nnz_heavy >= kSlabMinNnz && nnz_light >= (u64)8 * kNumSMs * kSpmvTile && nslabs16 >= 2) {
            const CsrView heavy{H, X.cols, nnz_heavy, X.off.get(), X.idx.get(), X.val.get()};
            build_range(heavy, X.chk, 0, false, true, s, st);
            DevBuf<u32> off2(X.rows - H + 1);
            k_rebase_offsets<<<grid_for(X.rows - H + 1, 256), 256, 0, s>>>(X.off, (u32)H, (u32)(X.rows - H), off2);
            LAS2_LAUNCH_CHECK();
            const CsrView light{X.rows - H, X.cols, nnz_light, off2.get(), X.idx.get() + nnz_heavy, X.val.get() + nnz_heavy};
            build_range(light, X.chk2, 0, false, false, s, st);
            X.chk2.row0 = H;
            LAS2_CUDA(cudaStreamSynchronize(s));
            return;
        }
    }
    build_range(whole, X.chk, X.hot_entries, slab_candidate, slab_mode == '1', s, st);
}

namespace {

void run_stream(const DeviceCsr &A, const ChunkStream &K, const double *x, double *y, cudaStream_t s, u64 *launches) {
    u32 hot = 0;
    if (K.slabs && K.slab_local) {
        hot = kSlabWidth;
    } else if (!K.slabs) {
        hot = A.tune.hot_entries >= 0 ? (u32)A.tune.hot_entries : std::min<u32>(K.hot_entries, kChunkHotDefault);
        hot = std::min<u32>(hot, kChunkHotMax);
        hot = (u32)std::min<u64>(hot, K.cols);
    }
    const size_t smem = std::min<size_t>(chunk_smem_bytes(hot) + (size_t)A.tune.smem_pad_kb * 1024, 227 * 1024);
    // The shared-memory carve-out is set explicitly to what this launch needs: left to itself the driver picks the
    // carve-out that would fit the most CTAs per SM (e.g. 196 KB for a 98 KB kernel), and the gathers then lose the
    // L1 lines their misses land in (scripts/micro/gather_l1_sweep.cu: 0.95 gathers/clk/SM with <= 132 KB carved
    // out, 0.43 with 228 KB).
    const int carve = A.tune.carveout_pct >= 0 ? A.tune.carveout_pct : (int)std::min<size_t>(100, (smem + 1024) * 100 / (228 * 1024) + 1);
    const int dev_slot = current_device_slot();
    auto configure = [&](const void *fn, int &cached) {
        if (cached == carve + 1) return;
        LAS2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        LAS2_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        cached = carve + 1;
    };
    // launch shape: 32 warps x 1 stage unless the knob "spmv_shape" says 16 x 2 (value 16)
    const bool wide = A.tune.shape != 16;
    const bool pdl = chunk_pdl_enabled(A.tune.pdl);
    const u32 warps = wide ? 32 : 16;
    const u32 grid = std::min<u32>(kNumSMs, cdiv(K.num_chunks, warps));
    const u32 *rowmap = (!K.slabs && K.has_empty_rows) ? K.rowmap.get() : nullptr;
    double *out = K.slabs ? K.part.get() : y;
    const int mode = (K.slabs && K.slab_local) ? 2 : (hot > 0 ? 1 : 0);
    if (!K.slabs && K.has_empty_rows) LAS2_CUDA(cudaMemsetAsync(y, 0, K.rows * sizeof(double), s));
#define LAS2_CHUNK_LAUNCH(MODE, WARPS, STAGES, SLOT)                                                                   \
    do {                                                                                                               \
        configure((const void *)k_spmv_chunk<MODE, WARPS, STAGES>, g_chunk_carve[dev_slot][SLOT]);                               \
        launch_pdl(pdl, k_spmv_chunk<MODE, WARPS, STAGES>, dim3(grid), dim3(WARPS * 32), smem, s, (const uint4 *)K.data,       \
                   (const u32 *)K.chunk_row, rowmap, x, out, (double *)K.carry, (double *)K.head, K.num_chunks,          \
                   mode == 2 ? 0u : hot, (const u32 *)K.slab_chunk_start, K.nslabs, (u32)K.cols);                      \
    } while (0)
    if (mode == 2) { if (wide) LAS2_CHUNK_LAUNCH(2, 32, 1, 0); else LAS2_CHUNK_LAUNCH(2, 16, 2, 1); }
    else if (mode == 1) { if (wide) LAS2_CHUNK_LAUNCH(1, 32, 1, 2); else LAS2_CHUNK_LAUNCH(1, 16, 2, 3); }
    else { if (wide) LAS2_CHUNK_LAUNCH(0, 32, 1, 4); else LAS2_CHUNK_LAUNCH(0, 16, 2, 5); }
#undef LAS2_CHUNK_LAUNCH
    launch_pdl(pdl, k_chunk_fixup, dim3(cdiv(K.num_chunks, 256)), dim3(256), 0, s, (const u32 *)K.fix_row, (const u32 *)K.fix_first,
               (const double *)K.carry, (const double *)K.head, out, K.num_chunks);
    if (launches) *launches += 2;
    if (K.slabs) {
        if (K.nslabs <= (u32)kFewSlabs)
            launch_pdl(pdl, k_slab_reduce_few, dim3(cdiv(K.rows, 256)), dim3(256), 0, s, (const u32 *)K.piece_id,
                       (const double *)K.part, (u64)K.rows, K.nslabs, y);
        else
            launch_pdl(pdl, k_slab_reduce, dim3(cdiv(K.rows, 32)), dim3(kReduceWarps * 32), 0, s, (const u32 *)K.piece_id,
                       (const double *)K.part, (u64)K.rows, K.nslabs, y);
        if (launches) *launches += 1;
    }
    LAS2_LAUNCH_CHECK();
}

} // namespace

void spmv_chunk(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches) {
    run_stream(A, A.chk, x, y + A.chk.row0, s, launches);
    if (A.chk2.active) run_stream(A, A.chk2, x, y + A.chk2.row0, s, launches);
}

// ---- host side of the 4-RHS product --------------------------------------------------------------------------------
bool spmm4_supported(const DeviceCsr &A) { return A.chk.active && !A.chk.slabs && !A.chk2.active; }

void Spmm4Work::ensure(const DeviceCsr &A) {
    const size_t nc = A.chk.num_chunks;
    if (carry4.size() < 4 * nc) { carry4.alloc(4 * nc); head4.alloc(4 * nc); }
    if (x4.size() < 4 * A.cols) x4.alloc(4 * A.cols);
    if (y4.size() < 4 * A.rows) y4.alloc(4 * A.rows);
    if (parts.size() < 4 * (size_t)kMaxParts) parts.alloc(4 * (size_t)kMaxParts);
}

// U[r] = A V[r] for r < nv <= 4 (V: vectors ld apart; U: vectors ldu apart) and the per-CTA partials of |U[r]|^2 in
// work.parts[r * kMaxParts ...] (nparts returned)
int spmm4(const DeviceCsr &A, const double *V, u64 ld, u32 nv, double *U, u64 ldu, Spmm4Work &work, cudaStream_t s, u64 *launches) {
    const ChunkStream &K = A.chk;
    work.ensure(A);
    const unsigned pgrid = std::max(1u, std::min<unsigned>(cdiv(A.cols, 256), kNumSMs * 8));
    k_pack4<<<pgrid, 256, 0, s>>>(V, ld, A.cols, nv, work.x4);
    const size_t smem = chunk_smem_bytes(0);
    static bool configured[kMaxDevicesTracked];
    const int dev_slot = current_device_slot();
    if (!configured[dev_slot]) {
        LAS2_CUDA(cudaFuncSetAttribute(k_spmm4_chunk<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        LAS2_CUDA(cudaFuncSetAttribute(k_spmm4_chunk<16, 2>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)std::min<size_t>(100, (smem + 1024) * 100 / (228 * 1024) + 1)));
        configured[dev_slot] = true;
    }
    const u32 *rowmap = K.has_empty_rows ? K.rowmap.get() : nullptr;
    if (K.has_empty_rows) LAS2_CUDA(cudaMemsetAsync(work.y4.get(), 0, 4 * K.rows * sizeof(double), s));
    const u32 grid = std::min<u32>(kNumSMs, cdiv(K.num_chunks, 16));
    k_spmm4_chunk<16, 2><<<grid, 16 * 32, smem, s>>>((const uint4 *)K.data, (const u32 *)K.chunk_row, rowmap, work.x4.get(),
                                                     work.y4.get(), work.carry4.get(), work.head4.get(), K.num_chunks);
    k_chunk_fixup4<<<cdiv(K.num_chunks, 256), 256, 0, s>>>((const u32 *)K.fix_row, (const u32 *)K.fix_first, work.carry4.get(),
                                                          work.head4.get(), work.y4.get(), K.num_chunks);
    u64 g = (A.rows + kVecThreads * 4 - 1) / (kVecThreads * 4);
    if (g < 1) g = 1;
    if (g > (u64)kMaxParts) g = kMaxParts;
    k_unpack4_sumsq<<<(unsigned)g, kVecThreads, 0, s>>>(work.y4.get(), A.rows, nv, U, ldu, work.parts.get(), (u32)kMaxParts);
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 4;
    return (int)g;
}

} // namespace las2
