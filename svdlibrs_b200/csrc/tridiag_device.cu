// tridiag_device.cu -- the O(js^3) half of imtql2 on the GPU.
//
// The reference updates all js rows of the eigenvector matrix z for every Givens rotation of the implicit-QL
// iteration (src/lib.rs:1647-1652, stride-js inner loop: ~0.5 s at js=800, minutes at js=3000 on one core).
// The rotation parameters depend only on the tridiagonal (d, e), never on z, so the host runs the cheap
// O(js^2) recurrences (host_math.cpp: ql_rotations) and this kernel replays the recorded rotations:
// one thread per row of z, rows independent, z stored transposed so that the threads of a warp touch
// consecutive addresses.  Within a sweep, rotation i touches columns (i, i+1) and i descends, so column i+1
// is final after rotation i and column i is carried in a register: one load + one store per rotation per
// row (16 B), the minimum.  Arithmetic is the reference's: two rounded products and one rounded add/sub.
#include "tridiag_device.cuh"

namespace las2 {

namespace {

__global__ void k_identity(double *__restrict__ zT, u32 ldz, u32 js) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 total = (u64)ldz * js;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        u32 col = (u32)(e / ldz), row = (u32)(e % ldz);
        zT[e] = (row == col) ? 1.0 : 0.0;
    }
}

constexpr int kReplayThreads = 64;

__global__ void __launch_bounds__(kReplayThreads)
k_ql_replay(double *__restrict__ zT, u32 ldz, u32 js, const QlSweep *__restrict__ sweeps, u32 nsweeps,
            const double2 *__restrict__ sc) {
    const u32 row = blockIdx.x * kReplayThreads + threadIdx.x;
    if (row >= js) return;
    double *z = zT + row;
    for (u32 w = 0; w < nsweeps; w++) {
        const QlSweep sw = sweeps[w];
        const double2 *__restrict__ rot = sc + sw.first;
        u32 i = sw.hi;
        double carry = z[(u64)(i + 1) * ldz];
#pragma unroll 4
        for (u32 t = 0; t < sw.count; t++, i--) {
            const double2 p = rot[t];                  // (s, c)
            const double x = z[(u64)i * ldz];
            z[(u64)(i + 1) * ldz] = __dadd_rn(__dmul_rn(p.x, x), __dmul_rn(p.y, carry));
            carry = __dsub_rn(__dmul_rn(p.y, x), __dmul_rn(p.x, carry));
        }
        z[(u64)(i + 1) * ldz] = carry; // i has been decremented once past the last rotation index
    }
}

__global__ void k_ritz_coefficients(const double *__restrict__ zT, u32 ldz, u32 js, const u32 *__restrict__ phys,
                                    u32 d, double *__restrict__ C) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (u64)js * d) return;
    u32 c = (u32)(e / js), i = (u32)(e % js); // consecutive threads walk down one column of z: coalesced reads
    C[(u64)i * d + c] = zT[(u64)phys[c] * ldz + (js - 1 - i)];
}

} // namespace

void ql_replay(const QlPlan &plan, u32 js, u32 ldz, double *zT, cudaStream_t s, u64 *launches, u64 *h2d_bytes) {
    unsigned g = cdiv((u64)ldz * js, 256);
    if (g > (unsigned)kNumSMs * 16) g = kNumSMs * 16;
    k_identity<<<g, 256, 0, s>>>(zT, ldz, js);
    if (launches) *launches += 1;
    if (!plan.sweeps.empty()) {
        DevBuf<QlSweep> dsw(plan.sweeps.size());
        DevBuf<double> dsc(plan.sc.size());
        LAS2_CUDA(cudaMemcpyAsync(dsw.get(), plan.sweeps.data(), plan.sweeps.size() * sizeof(QlSweep),
                                  cudaMemcpyHostToDevice, s));
        LAS2_CUDA(cudaMemcpyAsync(dsc.get(), plan.sc.data(), plan.sc.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        if (h2d_bytes) *h2d_bytes += plan.sweeps.size() * sizeof(QlSweep) + plan.sc.size() * sizeof(double);
        k_ql_replay<<<cdiv(js, kReplayThreads), kReplayThreads, 0, s>>>(zT, ldz, js, dsw, (u32)plan.sweeps.size(),
                                                                      reinterpret_cast<const double2 *>(dsc.get()));
        LAS2_LAUNCH_CHECK();
        if (launches) *launches += 1;
        LAS2_CUDA(cudaStreamSynchronize(s)); // dsw / dsc are freed on scope exit
    }
}

void ritz_coefficients(const double *zT, u32 ldz, u32 js, const u32 *phys_cols_dev, u32 d, double *C, cudaStream_t s,
                       u64 *launches) {
    if (d == 0) return;
    k_ritz_coefficients<<<cdiv((u64)js * d, 256), 256, 0, s>>>(zT, ldz, js, phys_cols_dev, d, C);
    LAS2_LAUNCH_CHECK();
    if (launches) *launches += 1;
}

} // namespace las2
