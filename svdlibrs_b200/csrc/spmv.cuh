// spmv.cuh -- sparse mat-vec entry points (see spmv.cu).
#pragma once
#include "device_csr.cuh"

namespace las2 {

extern int g_spmv_variant;     // -1 auto; test / tuning override (svdlas2_operator_set "spmv_variant")
extern int g_chunk_carveout_pct;  // chunk_spmv.cu: explicit carve-out (percent of 228 KB), -1 = exact fit
extern int g_chunk_pdl;           // chunk_spmv.cu: programmatic dependent launch on/off
extern int g_chunk_shape;         // chunk_spmv.cu: 32 (warps) x 1 (stage) or 16 x 2
extern int g_chunk_smem_pad_kb;   // chunk_spmv.cu: extra dynamic shared memory (L1 carve-out tuning)
extern int g_spmv_hot_entries; // -1 auto; entries of the gathered vector kept in shared memory ("spmv_hot")

// y = A x on stream s.  x: A.cols doubles, y: A.rows doubles, both device-resident, no aliasing.
void spmv(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches = nullptr);

} // namespace las2
