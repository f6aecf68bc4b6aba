// spmv.cuh -- sparse mat-vec entry points (see spmv.cu).
#pragma once
#include "device_csr.cuh"

namespace las2 {

// tuning knobs live in DeviceCsr::tune (per matrix, set through svdlas2_operator_set)

// y = A x on stream s.  x: A.cols doubles, y: A.rows doubles, both device-resident, no aliasing.
void spmv(const DeviceCsr &A, const double *x, double *y, cudaStream_t s, u64 *launches = nullptr);

// Four right-hand sides at once on the chunk stream (the sigma / U tail of ritvec): U[r] = A V[r], r < nv <= 4, plus the
// per-CTA partials of |U[r]|^2 in work.parts[r * kMaxParts + cta] (the number of partials is returned).  Every column is
// bit-identical to spmv() of that column.  Only for matrices whose chunk stream is in plain (not slab) order.
struct Spmm4Work {
    DevBuf<double> x4, y4, carry4, head4, parts;
    void ensure(const DeviceCsr &A);
};
bool spmm4_supported(const DeviceCsr &A);
int spmm4(const DeviceCsr &A, const double *V, u64 ld, u32 nv, double *U, u64 ldu, Spmm4Work &work, cudaStream_t s,
          u64 *launches = nullptr);

} // namespace las2
