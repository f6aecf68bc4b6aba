// device_csr.cu -- input upload: host CSR / CSC / COO (u64 indices, as nalgebra-sparse hands them out) ->
// device CSR(B) and CSR(B') with u32 indices, built entirely on the GPU and deterministically:
//   narrow + validate  ->  expand row ids  ->  stable LSD radix sort by the new major index
//   ->  offsets from key boundaries  ->  nnz-tile table for the streaming SpMV.
// Nothing here uses atomics on floating-point data; entries of every row end up in ascending minor order
// (for CSR/CSC inputs) or in stable triplet order among equal (row, col) (COO duplicates stay separate
// entries and are therefore "effectively summed" exactly like src/lib.rs:2121-2127).
#include "device_csr.cuh"

#include <algorithm>
#include <cstdlib>
#include <string>

#include "comm.hpp"
#include <vector>

namespace las2 {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems; // 4096 keys per CTA
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

// ------------------------------------------------------------------------------------------------ narrow

// dst[i] = (u32)src[i]; flags an error when src[i] > limit (offsets) or >= limit (indices).
__global__ void k_narrow(const u64 *__restrict__ src, u32 *__restrict__ dst, u64 n, u64 limit, int inclusive,
                         int *__restrict__ err) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        u64 v = src[i];
        bool bad = inclusive ? (v > limit) : (v >= limit);
        if (bad) *err = 1;
        dst[i] = (u32)v;
    }
}

__global__ void k_check_offsets(const u32 *__restrict__ off, u64 rows, u64 nnz, int *__restrict__ err) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    if (i == 0 && (off[0] != 0 || off[rows] != (u32)nnz)) *err = 2;
    for (; i < rows; i += stride)
        if (off[i] > off[i + 1]) *err = 2;
}

// rowid[k] = row containing nnz k  (upper_bound over the offsets, which stay L2-resident)
__global__ void k_expand_rows(const u32 *__restrict__ off, u32 rows, u64 nnz, u32 *__restrict__ rowid) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        u32 lo = 0, hi = rows; // first r with off[r+1] > k
        while (lo < hi) {
            u32 mid = lo + ((hi - lo) >> 1);
            if (off[mid + 1] > (u32)k) hi = mid; else lo = mid + 1;
        }
        rowid[k] = lo;
    }
}

// ------------------------------------------------------------------------------------------------ scan

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *total, u32 *smem /* >= 32 */) {
    // inclusive warp scan
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w; // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    u32 res = inc - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_local(u32 *__restrict__ data, u64 n, u32 *__restrict__ block_sums) {
    __shared__ u32 sm[33];
    u64 base = (u64)blockIdx.x * kScanChunk + (u64)threadIdx.x * kScanItems;
    u32 v[kScanItems];
    u32 sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        sum += v[i];
    }
    u32 total;
    u32 excl = block_exclusive_scan(sum, &total, sm);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) data[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(u32 *__restrict__ sums, u32 nb) {
    __shared__ u32 sm[33];
    u32 carry = 0;
    for (u32 base = 0; base < nb; base += kScanThreads) {
        u32 i = base + threadIdx.x;
        u32 v = (i < nb) ? sums[i] : 0;
        u32 total;
        u32 excl = block_exclusive_scan(v, &total, sm);
        if (i < nb) sums[i] = excl + carry;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(u32 *__restrict__ data, u64 n, const u32 *__restrict__ block_sums) {
    u32 add = block_sums[blockIdx.x];
    u64 base = (u64)blockIdx.x * kScanChunk + (u64)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
        if (base + i < n) data[base + i] += add;
}

// in-place exclusive scan of n u32 counters (sum < 2^32)
void exclusive_scan(u32 *data, u64 n, cudaStream_t s, UploadStats &st) {
    if (n == 0) return;
    u32 nb = cdiv(n, kScanChunk);
    DevBuf<u32> sums(nb);
    k_scan_local<<<nb, kScanThreads, 0, s>>>(data, n, sums);
    k_scan_sums<<<1, kScanThreads, 0, s>>>(sums, nb);
    k_scan_add<<<nb, kScanThreads, 0, s>>>(data, n, sums);
    LAS2_LAUNCH_CHECK();
    st.launches += 3;
    LAS2_CUDA(cudaStreamSynchronize(s)); // sums goes out of scope
}

// ------------------------------------------------------------------------------------------------ radix sort

// counts[bin * num_tiles + tile] = number of keys of the tile whose digit is bin
__global__ void __launch_bounds__(kSortThreads)
k_radix_hist(const u32 *__restrict__ keys, u64 n, int shift, u32 num_tiles, u32 *__restrict__ counts) {
    __shared__ u32 h[kRadix];
    h[threadIdx.x] = 0; // kSortThreads == kRadix
    __syncthreads();
    u64 base = (u64)blockIdx.x * kSortTile;
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        u64 k = base + (u64)i * kSortThreads + threadIdx.x;
        if (k < n) atomicAdd(&h[(keys[k] >> shift) & (kRadix - 1)], 1u); // integer counts: order-independent
    }
    __syncthreads();
    counts[(u64)threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  Warp w owns the contiguous 512 keys [w*512, (w+1)*512) of the tile and walks
// them in order, so equal digits keep their input order (this is what makes the transposed rows come out
// with ascending minor indices).
__global__ void __launch_bounds__(kSortThreads)
k_radix_scatter(const u32 *__restrict__ keys_in, const u32 *__restrict__ pay_in, const double *__restrict__ val_in,
                u32 *__restrict__ keys_out, u32 *__restrict__ pay_out, double *__restrict__ val_out, u64 n,
                int shift, u32 num_tiles, const u32 *__restrict__ bases) {
    __shared__ u32 wc[kSortWarps][kRadix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&wc[0][0])[i] = 0;
    __syncthreads();

    const u64 wbase = (u64)blockIdx.x * kSortTile + (u64)warp * (kSortItems * 32);
    u32 key[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        u64 k = wbase + (u64)i * 32 + lane;
        key[i] = (k < n) ? keys_in[k] : 0xffffffffu;
    }
    // phase 1: per-warp digit counts
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        u64 k = wbase + (u64)i * 32 + lane;
        bool valid = k < n;
        u32 d = (key[i] >> shift) & (kRadix - 1);
        u32 peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
        if (valid && lane == (__ffs(peers) - 1)) wc[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: digit d -> running base across the warps of this tile, seeded with the global base
    {
        u32 d = threadIdx.x;
        u32 run = bases[(u64)d * num_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            u32 c = wc[w][d];
            wc[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: ranked scatter, in order
#pragma unroll
    for (int i = 0; i < kSortItems; i++) {
        u64 k = wbase + (u64)i * 32 + lane;
        bool valid = k < n;
        u32 d = (key[i] >> shift) & (kRadix - 1);
        u32 peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
        u32 rank = __popc(peers & ((1u << lane) - 1u));
        u32 pos = 0;
        if (valid) pos = wc[warp][d] + rank;
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) wc[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) {
            keys_out[pos] = key[i];
            pay_out[pos] = pay_in[k];
            val_out[pos] = val_in[k];
        }
    }
}

struct SortBufs {
    DevBuf<u32> key[2], pay[2];
    DevBuf<double> val[2];
};

// Stable LSD radix sort of (key, pay, val) triples by the low `nbits` of key.  The first pass reads the
// caller's arrays; the result lands in bufs.*[which] with `which` returned.
int radix_sort(const u32 *key0, const u32 *pay0, const double *val0, u64 n, int nbits, SortBufs &bufs,
               cudaStream_t s, UploadStats &st) {
    u32 num_tiles = cdiv(n, kSortTile);
    DevBuf<u32> counts((u64)kRadix * num_tiles);
    for (int b = 0; b < 2; b++) {
        if (bufs.key[b].size() < n) { bufs.key[b].alloc(n); bufs.pay[b].alloc(n); bufs.val[b].alloc(n); }
    }
    int passes = (nbits + kRadixBits - 1) / kRadixBits;
    if (passes < 1) passes = 1;
    const u32 *kin = key0, *pin = pay0;
    const double *vin = val0;
    int dst = 0;
    for (int p = 0; p < passes; p++) {
        int shift = p * kRadixBits;
        k_radix_hist<<<num_tiles, kSortThreads, 0, s>>>(kin, n, shift, num_tiles, counts);
        st.launches += 1;
        exclusive_scan(counts, (u64)kRadix * num_tiles, s, st);
        k_radix_scatter<<<num_tiles, kSortThreads, 0, s>>>(kin, pin, vin, bufs.key[dst], bufs.pay[dst], bufs.val[dst],
                                                          n, shift, num_tiles, counts);
        LAS2_LAUNCH_CHECK();
        st.launches += 1;
        kin = bufs.key[dst]; pin = bufs.pay[dst]; vin = bufs.val[dst];
        dst ^= 1;
    }
    LAS2_CUDA(cudaStreamSynchronize(s));
    return dst ^ 1;
}

// off[c] = first position whose (sorted) key is >= c, for c in [0, nkeys]; derived from key boundaries.
__global__ void k_offsets_from_sorted(const u32 *__restrict__ keys, u64 n, u32 nkeys, u32 *__restrict__ off) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k <= n; k += stride) {
        u32 prev_next = (k == 0) ? 0 : keys[k - 1] + 1; // first key value not yet assigned
        u32 cur = (k == n) ? nkeys : keys[k];
        // all c in [prev_next, cur] start at k  (k == 0: c in [0, keys[0]])
        if (k == 0 || k == n || cur != keys[k - 1])
            for (u32 c = prev_next; c <= cur; c++) off[c] = (u32)k;
    }
}

__global__ void k_pad_tail(u32 *__restrict__ idx, double *__restrict__ val, u64 nnz, u64 padded) {
    u64 k = nnz + (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < padded) { idx[k] = 0; val[k] = 0.0; }
}

// tile_row[b] = first row r with off[r+1] > b*kSpmvTile ; tile_row[0] = 0 ; tile_row[num_tiles] = rows
__global__ void k_tile_rows(const u32 *__restrict__ off, u32 rows, u64 nnz, u32 num_tiles, u32 tile, u32 *__restrict__ tile_row) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > num_tiles) return;
    if (b == 0) { tile_row[0] = 0; return; }
    u64 start = (u64)b * tile;
    if (b == num_tiles || start >= nnz) { tile_row[b] = rows; return; }
    u32 lo = 0, hi = rows;
    while (lo < hi) {
        u32 mid = lo + ((hi - lo) >> 1);
        if (off[mid + 1] > (u32)start) hi = mid; else lo = mid + 1;
    }
    tile_row[b] = lo;
}

// number of column indices below `limit` (integer count: order-independent)
__global__ void k_count_below(const u32 *__restrict__ idx, u64 nnz, u32 limit, unsigned long long *__restrict__ out) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (; k < nnz; k += stride) c += idx[k] < limit;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// fix_row[b] = the row whose last piece is the head of tile b (it began in tile fix_first[b] < b), or ~0u
__global__ void k_fix_list(const u32 *__restrict__ off, const u32 *__restrict__ tile_row, u32 num_tiles, u32 tile,
                           u32 *__restrict__ fix_row, u32 *__restrict__ fix_first) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_tiles) return;
    u32 row = 0xffffffffu, first = 0;
    if (b > 0) {
        const u32 R0 = tile_row[b], R1 = tile_row[b + 1];
        if (R0 < R1) {
            const u32 o0 = off[R0];
            if (o0 < b * tile) { row = R0; first = o0 / tile; }
        }
    }
    fix_row[b] = row;
    fix_first[b] = first;
}

inline u64 pad_to_tile(u64 nnz) { return ((nnz + kSpmvTile - 1) / kSpmvTile) * kSpmvTile + kSpmvTile; }

void finish_csr(DeviceCsr &M, cudaStream_t s, UploadStats &st) {
    // pad, tile table, carry buffers
    u64 tail = M.padded_nnz - M.nnz;
    if (tail) {
        k_pad_tail<<<cdiv(tail, 256), 256, 0, s>>>(M.idx, M.val, M.nnz, M.padded_nnz);
        st.launches++;
    }
    M.num_tiles = cdiv(M.nnz, kSpmvTile);
    if (M.num_tiles == 0) M.num_tiles = 1; // an all-zero matrix still has rows to zero-fill
    M.tile_row.alloc(M.num_tiles + 1);
    M.carry.alloc(M.num_tiles);
    M.head.alloc(M.num_tiles);
    k_tile_rows<<<cdiv(M.num_tiles + 1, 256), 256, 0, s>>>(M.off, (u32)M.rows, M.nnz, M.num_tiles, kSpmvTile, M.tile_row);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    M.fix_row.alloc(M.num_tiles);
    M.fix_first.alloc(M.num_tiles);
    k_fix_list<<<cdiv(M.num_tiles, 256), 256, 0, s>>>(M.off, M.tile_row, M.num_tiles, kSpmvTile, M.fix_row, M.fix_first);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    // is a shared-memory prefix of the gathered vector worth it for this matrix?
    M.hot_entries = 0;
    M.hot_fraction = 0.0;
    if (M.nnz > 0) {
        const u32 limit = (u32)std::min<u64>(kHotPrefix, M.cols);
        DevBuf<unsigned long long> cnt(1);
        LAS2_CUDA(cudaMemsetAsync(cnt.get(), 0, sizeof(unsigned long long), s));
        unsigned grid = cdiv(M.nnz, 256 * 8);
        if (grid > (unsigned)kNumSMs * 8) grid = kNumSMs * 8;
        k_count_below<<<grid, 256, 0, s>>>(M.idx, M.nnz, limit, cnt);
        LAS2_LAUNCH_CHECK();
        st.launches++;
        unsigned long long h = 0;
        LAS2_CUDA(cudaMemcpyAsync(&h, cnt.get(), sizeof h, cudaMemcpyDeviceToHost, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
        M.hot_fraction = (double)h / (double)M.nnz;
        if (M.hot_fraction >= 0.25) M.hot_entries = limit;
    }
}

void check_err_flag(int *d_err, cudaStream_t s, const char *what) {
    int h = 0;
    LAS2_CUDA(cudaMemcpyAsync(&h, d_err, sizeof h, cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    if (h == 1) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, std::string(what) + ": index out of range");
    if (h == 2) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, std::string(what) + ": offsets are not a monotone prefix ending at nnz");
}

int bits_for(u64 n) { // bits needed to represent values in [0, n)
    int b = 1;
    while (b < 32 && ((u64)1 << b) < n) b++;
    return b;
}

// H2D of a u64 index array + narrowing to u32 on the device (halves nothing on the wire, but keeps the host
// pass-free; the wire is PCIe-bound either way).
void upload_narrow(const u64 *h, u64 n, u64 limit, int inclusive, u32 *d_dst, int *d_err, DevBuf<u64> &stage,
                   cudaStream_t s, UploadStats &st) {
    if (n == 0) return;
    if (stage.size() < n) stage.alloc(n);
    LAS2_CUDA(cudaMemcpyAsync(stage.get(), h, n * sizeof(u64), cudaMemcpyHostToDevice, s));
    st.h2d_bytes += n * sizeof(u64);
    unsigned grid = cdiv(n, 256 * 4);
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    k_narrow<<<grid, 256, 0, s>>>(stage.get(), d_dst, n, limit, inclusive, d_err);
    LAS2_LAUNCH_CHECK();
    st.launches++;
}

// CSR(X) resident on the device -> CSR(X') by one stable radix sort on the column index.
void transpose_csr(const DeviceCsr &X, DeviceCsr &XT, cudaStream_t s, UploadStats &st) {
    XT.rows = X.cols; XT.cols = X.rows; XT.nnz = X.nnz;
    XT.padded_nnz = pad_to_tile(XT.nnz);
    XT.off.alloc(XT.rows + 1);
    if (X.nnz == 0) {
        LAS2_CUDA(cudaMemsetAsync(XT.off.get(), 0, (XT.rows + 1) * sizeof(u32), s));
        XT.idx.alloc(XT.padded_nnz); XT.val.alloc(XT.padded_nnz);
        finish_csr(XT, s, st);
        return;
    }
    DevBuf<u32> rowid(X.nnz);
    {
        unsigned grid = cdiv(X.nnz, 256);
        if (grid > kNumSMs * 32) grid = kNumSMs * 32;
        k_expand_rows<<<grid, 256, 0, s>>>(X.off, (u32)X.rows, X.nnz, rowid);
        LAS2_LAUNCH_CHECK();
        st.launches++;
    }
    SortBufs bufs;
    // sort buffers sized to the padded length so the winner can be adopted as the CSR arrays directly
    for (int b = 0; b < 2; b++) { bufs.key[b].alloc(X.nnz); bufs.pay[b].alloc(XT.padded_nnz); bufs.val[b].alloc(XT.padded_nnz); }
    int w = radix_sort(X.idx, rowid, X.val, X.nnz, bits_for(X.cols), bufs, s, st);
    {
        unsigned grid = cdiv(X.nnz + 1, 256);
        if (grid > kNumSMs * 32) grid = kNumSMs * 32;
        k_offsets_from_sorted<<<grid, 256, 0, s>>>(bufs.key[w], X.nnz, (u32)XT.rows, XT.off);
        LAS2_LAUNCH_CHECK();
        st.launches++;
    }
    XT.idx = std::move(bufs.pay[w]);
    XT.val = std::move(bufs.val[w]);
    finish_csr(XT, s, st);
    LAS2_CUDA(cudaStreamSynchronize(s));
}

} // namespace

// helpers shared with chunk_spmv.cu
void exclusive_scan_u32(u32 *data, u64 n, cudaStream_t s, UploadStats &st) { exclusive_scan(data, n, s, st); }
void expand_row_ids(const u32 *off, u32 rows, u64 nnz, u32 *rowid, cudaStream_t s, UploadStats &st) {
    if (nnz == 0) return;
    unsigned grid = cdiv(nnz, 256);
    if (grid > (unsigned)kNumSMs * 32) grid = kNumSMs * 32;
    k_expand_rows<<<grid, 256, 0, s>>>(off, rows, nnz, rowid);
    LAS2_LAUNCH_CHECK();
    st.launches++;
}
void build_chunk_stream(DeviceCsr &X, cudaStream_t s, UploadStats &st); // chunk_spmv.cu

// the SpMV-ready copy of a finished CSR matrix: the warp-chunk stream (large matrices; small ones and matrices beyond
// its index range run the tile kernels on the natural CSR arrays)
static void build_streams(DeviceCsr &X, cudaStream_t s, UploadStats &st) { build_chunk_stream(X, s, st); }

// ------------------------------------------------------------------------------------------------ relabelling
namespace {

__global__ void k_row_lengths(const u32 *__restrict__ off, u32 rows, double *__restrict__ len) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; r < rows; r += stride) len[r] = (double)(off[r + 1] - off[r]);
}
// key = ~length (ascending key = descending length), payload = the row's old id, dummy value
__global__ void k_length_keys(const double *__restrict__ len, u32 rows, u32 *__restrict__ key, u32 *__restrict__ id,
                              double *__restrict__ dummy) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; r < rows; r += stride) {
        const double l = len[r];
        key[r] = 0xffffffffu - (u32)(l > 4294967295.0 ? 4294967295.0 : l);
        id[r] = r;
        dummy[r] = 0.0;
    }
}
__global__ void k_invert_perm(const u32 *__restrict__ new2old, u32 n, u32 *__restrict__ old2new) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; i < n; i += stride) old2new[new2old[i]] = i;
}
// lengths of the rows in their new order (ne[rows] = 0 so that the scan leaves the total there)
__global__ void k_permuted_lengths(const u32 *__restrict__ off, const u32 *__restrict__ new2old, u32 rows, u32 *__restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 stride = gridDim.x * blockDim.x;
    for (; i <= rows; i += stride) out[i] = i < rows ? off[new2old[i] + 1] - off[new2old[i]] : 0u;
}
// entries of new row i = entries of old row new2old[i], in their order
__global__ void k_gather_rows(const u32 *__restrict__ off_old, const u32 *__restrict__ idx_old, const double *__restrict__ val_old,
                              const u32 *__restrict__ off_new, const u32 *__restrict__ rowid_new, const u32 *__restrict__ new2old,
                              u64 nnz, u32 *__restrict__ idx_new, double *__restrict__ val_new) {
    u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; k < nnz; k += stride) {
        const u32 i = rowid_new[k];
        const u64 src = (u64)off_old[new2old[i]] + (k - off_new[i]);
        idx_new[k] = idx_old[src];
        val_new[k] = val_old[src];
    }
}

bool relabel_wanted(const DeviceCsr &P) {
    const char *e = std::getenv("SVDLAS2_RELABEL");
    if (e && e[0] == '0') return false;
    if (e && e[0] == '1') return P.nnz > 0 && P.rows > 1;
    // worth it where the chunk-stream kernels run and the gathered vector is longer than a shared-memory prefix
    return P.nnz >= (u64)8 * kNumSMs * kSpmvTile && P.rows > 2 * kHotPrefix;
}

} // namespace

// P = B' (rows = the N side).  Re-numbers its rows in descending order of GLOBAL length (stable), rebuilds P in that order and
// PT = B as its transpose (columns ascending in the new ids).  Deterministic; identical on every rank of a sharded run
// (the lengths are summed over the ranks first).
static void relabel_rows(DeviceCsr &P, DeviceCsr &PT, Relabel &R, Comm *comm, cudaStream_t s, UploadStats &st) {
    const u32 rows = (u32)P.rows;
    const unsigned g = std::max(1u, std::min<unsigned>(cdiv(rows, 256), kNumSMs * 8));
    DevBuf<double> len(rows);
    k_row_lengths<<<g, 256, 0, s>>>(P.off, rows, len);
    LAS2_LAUNCH_CHECK();
    if (comm) comm->allreduce_sum(len.get(), rows, s); // exact: integers far below 2^53
    DevBuf<u32> key(rows), id(rows);
    DevBuf<double> dummy(rows);
    k_length_keys<<<g, 256, 0, s>>>(len, rows, key, id, dummy);
    LAS2_LAUNCH_CHECK();
    st.launches += 2;
    SortBufs sb;
    const int w = radix_sort(key, id, dummy, rows, 32, sb, s, st);
    R.new2old = std::move(sb.pay[w]);
    R.old2new.alloc(rows);
    k_invert_perm<<<g, 256, 0, s>>>(R.new2old, rows, R.old2new);
    LAS2_LAUNCH_CHECK();
    R.h_new2old.resize(rows);
    LAS2_CUDA(cudaMemcpyAsync(R.h_new2old.data(), R.new2old.get(), rows * sizeof(u32), cudaMemcpyDeviceToHost, s));
    // P in the new row order
    DeviceCsr Pn;
    Pn.rows = P.rows; Pn.cols = P.cols; Pn.nnz = P.nnz; Pn.padded_nnz = P.padded_nnz; Pn.tune = P.tune;
    Pn.off.alloc(P.rows + 1);
    k_permuted_lengths<<<g, 256, 0, s>>>(P.off, R.new2old, rows, Pn.off);
    LAS2_LAUNCH_CHECK();
    exclusive_scan(Pn.off, (u64)rows + 1, s, st);
    Pn.idx.alloc(Pn.padded_nnz); Pn.val.alloc(Pn.padded_nnz);
    if (P.nnz) {
        DevBuf<u32> rowid(P.nnz);
        expand_row_ids(Pn.off, rows, P.nnz, rowid, s, st);
        unsigned gg = cdiv(P.nnz, 256 * 4);
        if (gg > (unsigned)kNumSMs * 32) gg = kNumSMs * 32;
        k_gather_rows<<<gg, 256, 0, s>>>(P.off, P.idx, P.val, Pn.off, rowid, R.new2old, P.nnz, Pn.idx, Pn.val);
        LAS2_LAUNCH_CHECK();
        st.launches++;
        LAS2_CUDA(cudaStreamSynchronize(s));
    }
    finish_csr(Pn, s, st);
    Pn.rows_sorted_desc = true;
    P = std::move(Pn);
    // PT = transpose of the re-ordered P: its column indices are the new ids, ascending inside every row
    const SpmvTuning keep = PT.tune;
    PT = DeviceCsr();
    transpose_csr(P, PT, s, st);
    PT.tune = keep;
    LAS2_CUDA(cudaStreamSynchronize(s));
    R.active = true;
}

// B / B' are complete in natural CSR form: relabel the N side where it pays, then build the SpMV streams
static void finalize_operator(DeviceCsr &B, DeviceCsr &BT, Relabel &R, Comm *comm, cudaStream_t s, UploadStats &st) {
    R = Relabel();
    if (relabel_wanted(BT)) relabel_rows(BT, B, R, comm, s, st);
    build_streams(B, s, st);
    build_streams(BT, s, st);
}

// host CSR arrays -> X (rows x cols) and XT in natural device CSR form (no streams yet)
static void upload_csr(u64 rows, u64 cols, u64 nnz, const u64 *offsets, const u64 *indices, const double *values,
                       cudaStream_t s, DeviceCsr &X, DeviceCsr &XT, UploadStats &st) {
    if (nnz >= ((u64)1 << 32) - 2 * kSpmvTile || rows >= ((u64)1 << 32) - 1 || cols >= ((u64)1 << 32) - 1)
        throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "matrix too large for 32-bit device indices (nnz, rows, cols must be < 2^32)");
    if (!offsets || (nnz && (!indices || !values))) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "null input array");
    DevBuf<int> err(1);
    LAS2_CUDA(cudaMemsetAsync(err.get(), 0, sizeof(int), s));
    X.rows = rows; X.cols = cols; X.nnz = nnz;
    X.padded_nnz = pad_to_tile(nnz);
    X.off.alloc(rows + 1);
    X.idx.alloc(X.padded_nnz);
    X.val.alloc(X.padded_nnz);
    DevBuf<u64> stage;
    upload_narrow(offsets, rows + 1, nnz, 1, X.off, err, stage, s, st);
    upload_narrow(indices, nnz, cols, 0, X.idx, err, stage, s, st);
    if (nnz) {
        LAS2_CUDA(cudaMemcpyAsync(X.val.get(), values, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
        st.h2d_bytes += nnz * sizeof(double);
    }
    k_check_offsets<<<cdiv(rows + 1, 256) > 2048 ? 2048 : cdiv(rows + 1, 256), 256, 0, s>>>(X.off, rows, nnz, err);
    LAS2_LAUNCH_CHECK();
    st.launches++;
    check_err_flag(err, s, "svdlas2 upload");
    stage.release();
    finish_csr(X, s, st);
    transpose_csr(X, XT, s, st);
}

void build_operator_from_csr(u64 rows, u64 cols, u64 nnz, const u64 *offsets, const u64 *indices, const double *values,
                             cudaStream_t s, DeviceCsr &X, DeviceCsr &XT, UploadStats &st, Relabel &relabel, Comm *comm) {
    upload_csr(rows, cols, nnz, offsets, indices, values, s, X, XT, st);
    finalize_operator(X, XT, relabel, comm, s, st);
}

void build_operator(const HostSparse &A, bool transposed, cudaStream_t s, DeviceCsr &B, DeviceCsr &BT, UploadStats &st,
                    Relabel &relabel) {
    DeviceCsr X, XT; // X = the matrix whose CSR the input arrays spell out; XT its transpose
    bool x_is_a = true;
    if (A.fmt == SVDLAS2_FMT_CSR) {
        upload_csr(A.nrows, A.ncols, A.nnz, A.a, A.b, A.val, s, X, XT, st);
    } else if (A.fmt == SVDLAS2_FMT_CSC) {
        // CSC(A) arrays are CSR(A')
        upload_csr(A.ncols, A.nrows, A.nnz, A.a, A.b, A.val, s, X, XT, st);
        x_is_a = false;
    } else if (A.fmt == SVDLAS2_FMT_COO) {
        if (A.nnz >= ((u64)1 << 32) - 2 * kSpmvTile || A.nrows >= ((u64)1 << 32) - 1 || A.ncols >= ((u64)1 << 32) - 1)
            throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "matrix too large for 32-bit device indices");
        if (A.nnz && (!A.a || !A.b || !A.val)) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "null input array");
        DevBuf<int> err(1);
        LAS2_CUDA(cudaMemsetAsync(err.get(), 0, sizeof(int), s));
        u64 nnz = A.nnz;
        X.rows = A.nrows; X.cols = A.ncols; X.nnz = nnz;
        X.padded_nnz = pad_to_tile(nnz);
        X.off.alloc(X.rows + 1);
        if (nnz == 0) {
            LAS2_CUDA(cudaMemsetAsync(X.off.get(), 0, (X.rows + 1) * sizeof(u32), s));
            X.idx.alloc(X.padded_nnz); X.val.alloc(X.padded_nnz);
        } else {
            DevBuf<u32> ri(nnz), ci(nnz);
            DevBuf<double> v(nnz);
            DevBuf<u64> stage;
            upload_narrow(A.a, nnz, A.nrows, 0, ri, err, stage, s, st);
            upload_narrow(A.b, nnz, A.ncols, 0, ci, err, stage, s, st);
            LAS2_CUDA(cudaMemcpyAsync(v.get(), A.val, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
            st.h2d_bytes += nnz * sizeof(double);
            check_err_flag(err, s, "svdlas2 upload (COO)");
            stage.release();
            // LSD: minor key (col) first, then major key (row); both stable => triplet order kept among duplicates
            SortBufs b1;
            int w1 = radix_sort(ci, ri, v, nnz, bits_for(A.ncols), b1, s, st); // key=col, pay=row
            SortBufs b2;
            for (int b = 0; b < 2; b++) { b2.key[b].alloc(nnz); b2.pay[b].alloc(X.padded_nnz); b2.val[b].alloc(X.padded_nnz); }
            int w2 = radix_sort(b1.pay[w1], b1.key[w1], b1.val[w1], nnz, bits_for(A.nrows), b2, s, st); // key=row, pay=col
            unsigned grid = cdiv(nnz + 1, 256);
            if (grid > kNumSMs * 32) grid = kNumSMs * 32;
            k_offsets_from_sorted<<<grid, 256, 0, s>>>(b2.key[w2], nnz, (u32)X.rows, X.off);
            LAS2_LAUNCH_CHECK();
            st.launches++;
            X.idx = std::move(b2.pay[w2]);
            X.val = std::move(b2.val[w2]);
            LAS2_CUDA(cudaStreamSynchronize(s));
        }
        finish_csr(X, s, st);
        transpose_csr(X, XT, s, st);
    } else {
        throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "unknown sparse format");
    }
    // B = A (not transposed) or A' (transposed)
    bool b_is_x = (x_is_a != transposed);
    if (b_is_x) { B = std::move(X); BT = std::move(XT); }
    else        { B = std::move(XT); BT = std::move(X); }
    finalize_operator(B, BT, relabel, nullptr, s, st);
    LAS2_CUDA(cudaStreamSynchronize(s));
}

} // namespace las2
