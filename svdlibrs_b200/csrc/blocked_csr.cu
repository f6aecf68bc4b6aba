// blocked_csr.cu -- builds the column-blocked copy of a device CSR matrix (see BlockedCsr in device_csr.cuh).
//
// Why this format exists: ncu on the plain-CSR kernels (profiles/r1a, r1b) shows the SpMV limited by the L1TEX
// tag stage -- one 128-byte line per cycle per SM for the uncoalesced x[col] gathers, i.e. ~3.4 ps per non-zero
// against 1.85 ps at the HBM roofline.  Shared memory serves random 8-byte reads ~5x faster, but only a
// 16 K-entry slice of the gathered vector fits next to the tile buffers.  So the non-zeros are regrouped by
// column block: while a CTA works on slab s, EVERY gather is an LDS into that slice.  The price is one partial
// sum (8 B written + 8 B read) and one offset (4 B) per (slab, row); a block is only turned into a slab when it
// holds enough non-zeros per row to pay for that (>= kMinPerRow on average), the rest goes to the remainder slab.
//
// Everything is built on the device and deterministically: integer counting, exclusive scan, ranked scatter.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_csr.cuh"

namespace las2 {

void exclusive_scan_u32(u32 *data, u64 n, cudaStream_t s, UploadStats &st); // device_csr.cu
void expand_row_ids(const u32 *off, u32 rows, u64 nnz, u32 *rowid, cudaStream_t s, UploadStats &st);
void build_tile_rows(const u32 *off, u32 rows, u64 nnz, u32 num_tiles, u32 tile, u32 *tile_row, cudaStream_t s, UploadStats &st);
void build_fix_list(const u32 *off, const u32 *tile_row, u32 num_tiles, u32 tile, u32 *fix_row, u32 *fix_first,
                    cudaStream_t s, UploadStats &st);

namespace {

constexpr double kMinPerRow = 1.7; // break-even of (10 + 20/len) B against the ~22 B/nnz-equivalent L1TEX path

// per-block non-zero counts (per-CTA shared histogram, then one integer atomic per bin per CTA)
__global__ void k_block_hist(const u32 *__restrict__ idx, u64 nnz, u32 nblocks, unsigned long long *__restrict__ out) {
    extern __shared__ u32 h[];
    for (u32 i = threadIdx.x; i < nblocks; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) atomicAdd(&h[idx[k] / kBlockWidth], 1u);
    __syncthreads();
    for (u32 i = threadIdx.x; i < nblocks; i += blockDim.x)
        if (h[i]) atomicAdd(&out[i], (unsigned long long)h[i]);
}

__device__ __forceinline__ u32 slab_of(u32 col, u32 nbk) {
    const u32 b = col / kBlockWidth;
    return b < nbk ? b : nbk;
}

// cnt[slab * rows + row] += 1 for every non-zero (integer atomics: order-independent)
__global__ void k_vrow_count(const u32 *__restrict__ idx, const u32 *__restrict__ rowid, u64 nnz, u32 nbk, u64 rows,
                             u32 *__restrict__ cnt) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) atomicAdd(&cnt[(u64)slab_of(idx[k], nbk) * rows + rowid[k]], 1u);
}

// voff[vr] += shift[slab(vr)] ; voff[vrows] = total
__global__ void k_shift_slabs(u32 *__restrict__ voff, u64 vrows, u64 rows, const u32 *__restrict__ shift, u32 total) {
    u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; v <= vrows; v += stride) voff[v] = (v == vrows) ? total : voff[v] + shift[v / rows];
}

// ranked scatter into the slab-ordered stream: inside a row the entries of one slab are contiguous (columns are
// ascending), so the rank of entry k is k minus the position of the slab's first entry in that row.
__global__ void k_blocked_scatter(const u32 *__restrict__ off, const u32 *__restrict__ idx, const double *__restrict__ val,
                                  const u32 *__restrict__ rowid, u64 nnz, u32 nbk, u64 rows,
                                  const u32 *__restrict__ voff, u64 rem_start, uint16_t *__restrict__ idx16,
                                  u32 *__restrict__ idx32, double *__restrict__ oval) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        const u32 c = idx[k], r = rowid[k];
        const u32 s = slab_of(c, nbk);
        const u32 first_col = s * kBlockWidth; // first column of the slab (the remainder slab starts at nbk * W)
        u32 lo = off[r], hi = (u32)k;          // lower_bound of first_col in idx[off[r] .. k]
        while (lo < hi) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (idx[mid] < first_col) lo = mid + 1; else hi = mid;
        }
        const u64 logical = (u64)voff[(u64)s * rows + r] + ((u32)k - lo);
        // interleaved tile layout: element j = t*8 + e of a tile lives at slot e*128 + t (see k_spmv_blocked)
        const u64 j = logical % kBlkTile;
        const u64 dest = logical - j + (j % 8) * (kBlkTile / 8) + j / 8;
        oval[dest] = val[k];
        if (s < nbk) idx16[dest] = (uint16_t)(c - first_col);
        else idx32[dest - rem_start] = c;
    }
}

inline u64 round_up_tile(u64 v) { return (v + kBlkTile - 1) / kBlkTile * kBlkTile; }

} // namespace

// Decides whether (and how far) to block `X` and builds X.blk.  COO-born matrices may hold duplicate columns in
// a row: still fine (duplicates are adjacent and stay in the same slab).
void build_blocked(DeviceCsr &X, cudaStream_t s, UploadStats &st) {
    BlockedCsr &K = X.blk;
    K = BlockedCsr();
    // Modes (env SVDLAS2_BLOCKED): unset = auto: the interleaved 32-bit stream WITHOUT slabs for matrices large
    // enough to fill the persistent kernel (profiles/r1_spmv_iterations.md: fastest on B200 so far); 0 = never;
    // 1 = slabs where they pay (experimental: the per-(slab,row) bookkeeping costs more than it saves at LSA
    // densities); 2 = no slabs, always.
    const char *want = std::getenv("SVDLAS2_BLOCKED");
    const char mode = (want && want[0]) ? want[0] : 'a';
    if (mode == '0') return;
    if (mode == 'a' && X.nnz < (u64)8 * kNumSMs * kSpmvTile) return;
    const bool no_slabs = mode != '1';
    if (X.nnz == 0 || X.rows == 0) return;
    const u32 nblocks = cdiv(X.cols, kBlockWidth);
    if (nblocks > 12000) return; // shared histogram limit; such matrices are far too sparse per block anyway

    // 1. non-zeros per column block
    DevBuf<unsigned long long> dh(nblocks);
    LAS2_CUDA(cudaMemsetAsync(dh.get(), 0, nblocks * sizeof(unsigned long long), s));
    {
        unsigned grid = cdiv(X.nnz, 256 * 16);
        if (grid > (unsigned)kNumSMs * 8) grid = kNumSMs * 8;
        k_block_hist<<<grid, 256, nblocks * sizeof(u32), s>>>(X.idx, X.nnz, nblocks, dh);
        LAS2_LAUNCH_CHECK();
        st.launches++;
    }
    std::vector<unsigned long long> hist(nblocks);
    LAS2_CUDA(cudaMemcpyAsync(hist.data(), dh.get(), nblocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));

    // 2. longest prefix of blocks that pays: include block b while it holds >= kMinPerRow entries per row
    u32 nbk = 0;
    u64 in_slabs = 0;
    while (!no_slabs && nbk < nblocks && (double)hist[nbk] >= kMinPerRow * (double)X.rows) { in_slabs += hist[nbk]; nbk++; }
    // a long, nearly uniform tail (B' of an LSA matrix) oscillates around the threshold: judge it as a whole
    if (!no_slabs && nbk < nblocks) {
        u64 rest = X.nnz - in_slabs;
        if ((double)rest >= kMinPerRow * (double)X.rows * (double)(nblocks - nbk)) { nbk = nblocks; in_slabs = X.nnz; }
    }
    if (nbk == 0 && !no_slabs) return;
    const bool has_rem = in_slabs < X.nnz;
    const u32 nslabs = nbk + (has_rem ? 1 : 0);
    const u64 vrows = (u64)nslabs * X.rows;
    if (vrows >= ((u64)1 << 31)) return;

    // 3. entries per virtual row -> positions
    DevBuf<u32> rowid(X.nnz);
    expand_row_ids(X.off, (u32)X.rows, X.nnz, rowid, s, st);
    K.voff.alloc(vrows + 1);
    LAS2_CUDA(cudaMemsetAsync(K.voff.get(), 0, (vrows + 1) * sizeof(u32), s));
    if (nbk == 0) {
        // one slab holding everything: the positions are the row offsets themselves
        LAS2_CUDA(cudaMemcpyAsync(K.voff.get(), X.off.get(), (X.rows + 1) * sizeof(u32), cudaMemcpyDeviceToDevice, s));
    } else {
        unsigned grid = cdiv(X.nnz, 256 * 4);
        if (grid > (unsigned)kNumSMs * 16) grid = kNumSMs * 16;
        k_vrow_count<<<grid, 256, 0, s>>>(X.idx, rowid, X.nnz, nbk, X.rows, K.voff);
        LAS2_LAUNCH_CHECK();
        st.launches++;
        exclusive_scan_u32(K.voff, vrows + 1, s, st);
    }
    // slab starts (unpadded) -> tile-padded starts
    std::vector<u32> starts(nslabs + 1);
    for (u32 i = 0; i <= nslabs; i++)
        LAS2_CUDA(cudaMemcpyAsync(&starts[i], K.voff.get() + (u64)i * X.rows, sizeof(u32), cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    std::vector<u32> shift(nslabs), tile_start(nslabs + 1);
    u64 pos = 0;
    for (u32 i = 0; i < nslabs; i++) {
        shift[i] = (u32)(pos - starts[i]);
        tile_start[i] = (u32)(pos / kBlkTile);
        pos = round_up_tile(pos + (starts[i + 1] - starts[i]));
    }
    if (pos + kBlkTile >= ((u64)1 << 32)) { K = BlockedCsr(); return; }
    tile_start[nslabs] = (u32)(pos / kBlkTile);
    K.padded = pos;
    K.rem_start = has_rem ? (u64)tile_start[nbk] * kBlkTile : pos;
    DevBuf<u32> dshift(nslabs);
    LAS2_CUDA(cudaMemcpyAsync(dshift.get(), shift.data(), nslabs * sizeof(u32), cudaMemcpyHostToDevice, s));
    {
        unsigned grid = cdiv(vrows + 1, 256 * 4);
        if (grid > (unsigned)kNumSMs * 16) grid = kNumSMs * 16;
        k_shift_slabs<<<grid, 256, 0, s>>>(K.voff, vrows, X.rows, dshift, (u32)pos);
        LAS2_LAUNCH_CHECK();
        st.launches++;
    }

    // 4. the stream itself (+ one spare tile so that vector loads past the end stay in bounds)
    K.val.alloc(pos + kBlkTile);
    K.idx16.alloc(K.rem_start + kBlkTile);
    K.idx32.alloc(pos - K.rem_start + kBlkTile);
    LAS2_CUDA(cudaMemsetAsync(K.val.get(), 0, (pos + kBlkTile) * sizeof(double), s));
    LAS2_CUDA(cudaMemsetAsync(K.idx16.get(), 0, (K.rem_start + kBlkTile) * sizeof(uint16_t), s));
    LAS2_CUDA(cudaMemsetAsync(K.idx32.get(), 0, (pos - K.rem_start + kBlkTile) * sizeof(u32), s));
    {
        unsigned grid = cdiv(X.nnz, 256 * 2);
        if (grid > (unsigned)kNumSMs * 32) grid = kNumSMs * 32;
        k_blocked_scatter<<<grid, 256, 0, s>>>(X.off, X.idx, X.val, rowid, X.nnz, nbk, X.rows, K.voff, K.rem_start,
                                               K.idx16, K.idx32, K.val);
        LAS2_LAUNCH_CHECK();
        st.launches++;
    }

    // 5. tile table over virtual rows, carries, partial sums
    K.num_tiles = (u32)(pos / kBlkTile);
    K.blocked_tiles = (u32)(K.rem_start / kBlkTile);
    K.tile_row.alloc(K.num_tiles + 1);
    build_tile_rows(K.voff, (u32)vrows, pos, K.num_tiles, kBlkTile, K.tile_row, s, st);
    K.slab_tile_start.alloc(nslabs + 1);
    LAS2_CUDA(cudaMemcpyAsync(K.slab_tile_start.get(), tile_start.data(), (nslabs + 1) * sizeof(u32), cudaMemcpyHostToDevice, s));
    K.carry.alloc(K.num_tiles);
    K.head.alloc(K.num_tiles);
    K.fix_row.alloc(K.num_tiles);
    K.fix_first.alloc(K.num_tiles);
    build_fix_list(K.voff, K.tile_row, K.num_tiles, kBlkTile, K.fix_row, K.fix_first, s, st);
    if (nslabs > 1) K.partial.alloc(vrows);
    LAS2_CUDA(cudaStreamSynchronize(s)); // rowid, dshift, host vectors go out of scope

    K.nbk = nbk; K.nslabs = nslabs; K.has_rem = has_rem;
    K.rows = X.rows; K.cols = X.cols; K.vrows = vrows;
    K.stream_bytes = K.rem_start * 10 + (pos - K.rem_start) * 12 + (vrows + 1) * 4 + (nslabs > 1 ? vrows * 16 : 0) +
                     X.rows * 8 + X.cols * 8;
    K.active = true;
}

} // namespace las2
