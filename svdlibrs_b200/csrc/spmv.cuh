// spmv.cuh -- sparse mat-vec entry points (see spmv.cu).
#pragma once
#include "device_csr.cuh"

namespace las2 {

// tuning knobs live in DeviceCsr::tune (per matrix, set through svdlas2_operator_set)

// y = A x on stream s.  x: A.cols doubles, y: A.rows doubles, both device-resident, no aliasing.
void spmv(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches = nullptr);

} // namespace las2
