// Microbenchmark: does the rate of random 8-byte gathers (L2-resident vector) depend on how much of the SM's
// 256 KB is carved out as shared memory (i.e. on the L1 capacity left for lines in flight)?  One CTA of 1024
// threads per SM, 8 independent gathers per thread per iteration, dynamic shared memory S KB (unused).
// Load flavours: nc (ld.global.nc), na (ld.global.nc.L1::no_allocate), cg (ld.global.cg: L2 only), mix (every other
// gather from a 4096-entry shared-memory table instead).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_l1_sweep gather_l1_sweep.cu && ./gather_l1_sweep
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ double ldx(const double *p) {
    double r;
    if (MODE == 0) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    else asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

template <int MODE, bool MIX>
__global__ void __launch_bounds__(1024, 1) k_gather(const uint32_t *__restrict__ idx, const double *__restrict__ x, double *out, size_t n) {
    extern __shared__ double xs[];
    if (MIX) { for (int i = threadIdx.x; i < 4096; i += blockDim.x) xs[i] = x[i]; __syncthreads(); }
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 7 * stride < n; i += 8 * stride) {
        uint32_t c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = idx[i + k * stride];
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MIX && (k & 1)) v[k] = xs[c[k] & 4095u];
            else v[k] = ldx<MODE>(x + c[k]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) s += v[k];
    }
    if (s == 123.456) out[0] = s;
}

int main() {
    const size_t n = (size_t)1 << 27;
    for (uint32_t m : {1000000u}) {
        std::vector<uint32_t> h(n);
        uint64_t st = 88172645463325252ull;
        for (size_t i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i] = (uint32_t)(st % m); }
        uint32_t *idx; double *x, *out;
        CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&x, (size_t)m * 8)); CK(cudaMalloc(&out, 8));
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(x, 0, (size_t)m * 8));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto time = [&](auto f) { f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < 3; r++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 3; };
        for (int kb = 8; kb <= 224; kb += 8) {
            const int S = kb * 1024;
            CK(cudaFuncSetAttribute(k_gather<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S));
            CK(cudaFuncSetAttribute(k_gather<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S));
            CK(cudaFuncSetAttribute(k_gather<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S));
            CK(cudaFuncSetAttribute(k_gather<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S));
            float a = time([&] { k_gather<0, false><<<148, 1024, S>>>(idx, x, out, n); });
            float b = time([&] { k_gather<1, false><<<148, 1024, S>>>(idx, x, out, n); });
            float c = time([&] { k_gather<2, false><<<148, 1024, S>>>(idx, x, out, n); });
            float d = time([&] { k_gather<0, true><<<148, 1024, S>>>(idx, x, out, n); });
            CK(cudaGetLastError());
            auto r = [&](float ms) { return n / ms / 1e6 / 148 / 1.965; };
            printf("m=%8u smem=%3d KB  gathers/clk/SM: nc %.2f | no_allocate %.2f | cg %.2f | mix(half LDS) %.2f  (total elements/clk/SM)\n", m, kb, r(a), r(b), r(c), r(d));
        }
        cudaFree(idx); cudaFree(x); cudaFree(out);
    }
    return 0;
}
