// device_csr.cuh -- the operator as it lives in HBM: CSR(B) and CSR(B') with 32-bit indices.
//
// Reference: the matrix is only ever touched through SMat::svd_opa (src/lib.rs:2035-2130), whose
// `transposed` branch is a scatter loop.  Here the input is converted ONCE at upload into row-gather form
// for both B and B' (north star "Input upload"), so neither half of svd_opb (:829-837) needs atomics.
#pragma once

#include <vector>

#include "util.cuh"

namespace las2 {

// nnz-tile size of the streaming SpMV (see spmv.cu); value/index arrays are zero-padded to a multiple of it.
constexpr u32 kSpmvTile = 2048;
// candidate size of the shared-memory prefix of the gathered vector (64 KB of doubles)
constexpr u32 kHotPrefix = 8192;

// Warp-chunk stream (chunk_spmv.cu): the row-ordered non-zeros cut into chunks of 256, one warp per chunk.  A
// chunk is 3072 contiguous bytes, fetched by ONE bulk copy: 4 x 32 x 16 B of values, then 2 x 32 x 16 B of
// column indices, arranged so that six conflict-free 128-bit shared-memory reads hand lane l the 8 CONSECUTIVE
// elements l*8 .. l*8+7 of the chunk.  Bit 31 of an index marks the first element of a row.
constexpr u32 kChunkE = 8;                    // consecutive elements per lane
constexpr u32 kChunkNnz = 32 * kChunkE;       // 256
constexpr u32 kChunkBytes = kChunkNnz * 12;   // 3072
constexpr u32 kChunkFlag = 0x80000000u;

// Slab mode (matrices with enough entries per (column block, row) piece): the same chunk stream, but ordered
// (slab, row, column) with slab = column / kSlabWidth and LOCAL column indices, so that while a CTA works inside a slab
// every gather is a shared-memory read of the slab's 128 KB slice of x.  A "row" of the stream is then a non-empty
// (slab, row) piece; the kernel writes one partial sum per piece (indexed by the piece's running number) and
// k_slab_reduce adds the pieces of each row in slab order.
constexpr u32 kSlabWidth = 16384;

struct ChunkStream {
    bool active = false;
    bool slabs = false;
    // slab flavours: LOCAL (width kSlabWidth, slab-local column indices, every gather from the slab's slice of x in shared
    // memory) or WIDE (any width, global column indices, gathers through L1: the stream is merely ORDERED by slabs so that all
    // CTAs gather from the same few MB of x at the same time -- an L2 / TLB blocking for gathered vectors far larger than L2's
    // useful share, e.g. the 80 MB of t at cfg 4)
    bool slab_local = false;
    u32 slab_width = 0;
    u64 row0 = 0;            // first row of y this stream produces (a matrix may be split into row ranges with different modes)
    u32 hot_entries = 0;     // plain mode: size of the shared-memory prefix of x worth keeping (0: none)
    u64 rows = 0, cols = 0;
    u32 num_chunks = 0;
    u32 nonempty_rows = 0;   // plain: rows with entries; slab mode: non-empty (slab, row) pieces
    bool has_empty_rows = false;
    u32 nslabs = 0;
    DevBuf<uint4> data;      // (num_chunks + 1) * 192
    DevBuf<u32> chunk_row;   // num_chunks + 1: index (among the non-empty rows) of the first row that STARTS at or after
                             // the chunk's first element; bit 31: some lane of the chunk holds two or more row starts
    DevBuf<u32> rowmap;      // non-empty row index -> row (plain mode only, empty when every row is non-empty)
    DevBuf<double> carry;    // num_chunks: sum of the piece that is still open at the chunk's end
    DevBuf<double> head;     // num_chunks: sum of the piece that was open at the chunk's start, up to the first row start
    DevBuf<u32> fix_row, fix_first; // num_chunks: row completed by chunk b's head (or ~0u) and the chunk it started in
    // slab mode
    DevBuf<u32> slab_chunk_start;   // nslabs + 1
    DevBuf<u32> piece_id;           // nslabs * rows: running number of piece (slab, row), 0xffffffff when it has no entries
    DevBuf<double> part;            // one partial sum per piece
    u64 stream_bytes = 0;           // bytes one SpMV moves with this format (reporting)
};

// Per-matrix SpMV tuning knobs (svdlas2_operator_set "spmv_*"): they belong to the operator they were set on, so two
// operators of one process never see each other's settings (the driver is re-entrant).
struct SpmvTuning {
    int variant = -1;      // -1 auto; 0/1/2 tile kernels with 256/128/512 threads, 3 persistent hot-prefix tile kernel, 5 chunk stream
    int hot_entries = -1;  // -1 auto; entries of the gathered vector kept in shared memory
    int carveout_pct = -1; // chunk kernel: shared-memory carve-out in percent of 228 KB (-1: exact fit)
    int pdl = -1;          // chunk kernel chain: programmatic dependent launch (-1: SVDLAS2_PDL, default on)
    int shape = 32;        // chunk kernel: 32 warps x 1 stage, or 16 warps x 2 stages (value 16)
    int smem_pad_kb = 0;   // chunk kernel: extra (unused) dynamic shared memory
};

struct DeviceCsr {
    u64 rows = 0, cols = 0, nnz = 0;
    SpmvTuning tune;
    ChunkStream chk;  // rows [0, chk2.row0) -- or all rows when chk2 is not active
    ChunkStream chk2; // optional second row range (the light rows of a frequency-ordered B')
    // set by the upload-time relabelling for the matrix whose ROWS were put in descending order of (global) row length
    bool rows_sorted_desc = false;
    DevBuf<u32> off;    // rows + 1 row offsets
    DevBuf<u32> idx;    // padded_nnz column indices (padding: index 0)
    DevBuf<double> val; // padded_nnz values        (padding: 0.0)
    u64 padded_nnz = 0;
    // streaming-SpMV tile table: tile b covers nnz [b*kSpmvTile, (b+1)*kSpmvTile)
    u32 num_tiles = 0;
    DevBuf<u32> tile_row;  // num_tiles + 1: first row that ENDS after the tile's first nnz (tile_row[0] = 0)
    DevBuf<double> carry;  // num_tiles: partial sum of the row that continues into the next tile
    DevBuf<double> head;   // num_tiles: partial sum of the tile's first row when it began in an earlier tile
    DevBuf<u32> fix_row, fix_first; // num_tiles: row completed by the fix-up of tile b (or ~0u) and its first tile
    // hot prefix of the gathered vector: entries [0, hot_entries) are worth keeping in shared memory when a
    // large share of the column indices falls there (measured at upload: hot_fraction)
    u32 hot_entries = 0;
    double hot_fraction = 0.0;
    // algorithmic bytes of one y = A x with this matrix: 12*Z + 4*(R+1) + 8*C + 8*R
    u64 spmv_bytes() const { return 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows; }
};

struct HostSparse {
    int fmt; // SVDLAS2_FMT_*
    u64 nrows, ncols, nnz;
    const u64 *a; // CSR/CSC: major offsets; COO: row indices
    const u64 *b; // CSR/CSC: minor indices; COO: col indices
    const double *val;
};

struct UploadStats {
    u64 h2d_bytes = 0;
    u64 launches = 0;
};

// Upload-time relabelling of the N side (the columns of B = the rows of B'), SURVEY 8(a) svd_opa: ids are re-assigned in
// descending order of column popularity (stable), so that
//   * the popular entries of the gathered vector x form a PREFIX whatever order the caller's matrix uses (the shared-memory
//     prefix of k_spmv_chunk then serves half of the gathers of an LSA-shaped matrix), and
//   * the rows of B' come out sorted by length, so that its heavy rows can run in slab mode and its light rows apart.
// The Lanczos recurrence runs on P B'B P' (same mathematics, vectors in the new order); the start vector is permuted on the
// way in and the N-side singular vectors on the way out, so callers never see the new ids.
struct Relabel {
    bool active = false;
    DevBuf<u32> new2old, old2new;   // device copies (N entries each)
    std::vector<u32> h_new2old;     // host copy (start vector, test hooks)
};

class Comm;

// Builds CSR(B) and CSR(B') on the current device.  `transposed` is the orientation decision of
// src/lib.rs:606 (B = A' when true).  Throws Error(SVDLAS2_ERR_INVALID_ARGUMENT) on malformed input.
void build_operator(const HostSparse &A, bool transposed, cudaStream_t stream, DeviceCsr &B, DeviceCsr &BT,
                    UploadStats &stats, Relabel &relabel);

// Builds the pair from a CSR row shard that is already oriented (multi-GPU path).
void build_operator_from_csr(u64 rows, u64 cols, u64 nnz, const u64 *offsets, const u64 *indices,
                             const double *values, cudaStream_t stream, DeviceCsr &B, DeviceCsr &BT,
                             UploadStats &stats, Relabel &relabel, Comm *comm);

} // namespace las2
