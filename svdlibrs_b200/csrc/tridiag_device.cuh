// tridiag_device.cuh -- GPU replay of the imtql2 eigenvector rotations (reference src/lib.rs:1646-1652) and
// assembly of the Ritz coefficient matrix used by ritvec (:1807-1817).
#pragma once
#include "host_math.hpp"
#include "util.cuh"

namespace las2 {

// zT (device): js columns x ldz rows, zT[col * ldz + row] = z[row][col] of the reference's row-major z.
// Sets zT = I, uploads the plan and replays every rotation; one thread per row of z (rows are independent).
extern int g_ql_replay_variant; // test hook: -1 default, 0 one replaying warp per CTA, 1 pipeline of sweeps over 8 warps
void ql_replay(const QlPlan &plan, u32 js, u32 ldz, double *zT, cudaStream_t s, u64 *launches, u64 *h2d_bytes);

// C[i * d + c] = z[(js-1-i)][logical column k_c] = zT[phys_cols[c] * ldz + (js-1-i)]   (T was reversed, :1790-1791)
void ritz_coefficients(const double *zT, u32 ldz, u32 js, const u32 *phys_cols_dev, u32 d, double *C, cudaStream_t s,
                       u64 *launches);

} // namespace las2
