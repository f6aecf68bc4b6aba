// Microbenchmark: throughput of random 8-byte gathers from an L2-resident vector on B200:
//   (a) ld.global.nc (LSU / L1TEX tag stage), (b) tex1Dfetch<int2> (texture path), (c) shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_ldg(const uint32_t *__restrict__ idx, const double *__restrict__ x, double *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint32_t a = idx[i], b = idx[i + stride], c = idx[i + 2 * stride], d = idx[i + 3 * stride];
        s += __ldg(x + a) + __ldg(x + b) + __ldg(x + c) + __ldg(x + d);
    }
    if (s == 123.456) out[0] = s;
}
__global__ void k_tex(const uint32_t *__restrict__ idx, cudaTextureObject_t tx, double *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint32_t a = idx[i], b = idx[i + stride], c = idx[i + 2 * stride], d = idx[i + 3 * stride];
        int2 va = tex1Dfetch<int2>(tx, a), vb = tex1Dfetch<int2>(tx, b), vc = tex1Dfetch<int2>(tx, c), vd = tex1Dfetch<int2>(tx, d);
        s += __hiloint2double(va.y, va.x) + __hiloint2double(vb.y, vb.x) + __hiloint2double(vc.y, vc.x) + __hiloint2double(vd.y, vd.x);
    }
    if (s == 123.456) out[0] = s;
}
__global__ void k_smem(const uint32_t *__restrict__ idx, const double *__restrict__ x, double *out, size_t n, uint32_t m) {
    extern __shared__ double xs[];
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) xs[i] = x[i];
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    double s = 0;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint32_t a = idx[i] % m, b = idx[i + stride] % m, c = idx[i + 2 * stride] % m, d = idx[i + 3 * stride] % m;
        s += xs[a] + xs[b] + xs[c] + xs[d];
    }
    if (s == 123.456) out[0] = s;
}
int main() {
    const size_t n = (size_t)1 << 27; // 134 M gathers
    for (uint32_t m : {100000u, 1000000u, 10000000u}) {
        std::vector<uint32_t> h(n);
        uint64_t st = 88172645463325252ull;
        for (size_t i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i] = (uint32_t)(st % m); }
        uint32_t *idx; double *x, *out;
        CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&x, (size_t)m * 8)); CK(cudaMalloc(&out, 8));
        CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemset(x, 0, (size_t)m * 8));
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = x;
        rd.res.linear.desc = cudaCreateChannelDesc(32, 32, 0, 0, cudaChannelFormatKindSigned); rd.res.linear.sizeInBytes = (size_t)m * 8;
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tx; CK(cudaCreateTextureObject(&tx, &rd, &td, nullptr));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto time = [&](auto f) { f(); cudaDeviceSynchronize(); cudaEventRecord(e0); for (int r = 0; r < 5; r++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5; };
        for (int tpb : {256, 1024}) {
            int grid = 148 * (2048 / tpb);
            float a = time([&] { k_ldg<<<grid, tpb>>>(idx, x, out, n); });
            float b = time([&] { k_tex<<<grid, tpb>>>(idx, tx, out, n); });
            printf("m=%8u tpb=%4d  ldg %.3f ms = %.1f G gathers/s (%.2f /clk/SM @1.965GHz) | tex %.3f ms = %.1f G/s (%.2f /clk/SM)\n", m, tpb,
                   a, n / a / 1e6, n / a / 1e6 / 148 / 1.965, b, n / b / 1e6, n / b / 1e6 / 148 / 1.965);
        }
        uint32_t ms_ = 16384;
        CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        float c = time([&] { k_smem<<<148, 1024, 131072>>>(idx, x, out, n, ms_); });
        printf("           smem(16384 entries, 1024 thr/SM) %.3f ms = %.1f G gathers/s (%.2f /clk/SM)\n", c, n / c / 1e6, n / c / 1e6 / 148 / 1.965);
        CK(cudaGetLastError());
        cudaDestroyTextureObject(tx); cudaFree(idx); cudaFree(x); cudaFree(out);
    }
    return 0;
}
