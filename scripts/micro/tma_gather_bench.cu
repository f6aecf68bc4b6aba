// Microbenchmark: can the TMA unit serve random 8-byte gathers NEXT TO the LSU path?  (round 2)
// The chunk SpMV is bound by the L1TEX tag stage: ~1 uncoalesced gather per clock per SM (gather_bench.cu).  The TMA
// unit reads L2 without going through L1TEX, so gathers issued as `cp.async.bulk.tensor ... tile::gather4` (four
// 16-byte rows of a [V/2 x 2] f64 tensor per instruction) or as 16-byte 1-D bulk copies could be a second, independent
// gather path.  Variants (one per process run: a faulting variant must not poison the others):
//   ldg      8 x ld.global.nc per lane and iteration                                (the L1TEX baseline)
//   g4       2 x gather4 per lane and iteration, nothing through the LSU            (TMA alone)
//   bulk     8 x 16-byte cp.async.bulk per lane and iteration                       (TMA alone, 1-D)
//   mix      1 x gather4 + 4 x ld.global.nc per lane and iteration                  (both paths at once)
//   mixb     4 x 16-byte cp.async.bulk + 4 x ld.global.nc per lane and iteration
// gather4 wants its 64-byte landing zone 128-byte aligned (dst stride 128); the 1-D copies only 16-byte aligned.
// Every variant sums what it gathered; the sums must agree (exact in f64 for the test vector).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_bench tma_gather_bench.cu
//   ./tma_gather_bench <variant> <vector doubles> [box rows = 1] [warps per CTA = 16] [stages = 4] [dst stride = 64]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_addr(dst)), "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk16(void *dst, const void *src, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(smem_addr(bar)) : "memory");
}

// idx is laid out so that (iteration, warp, lane) reads 8 consecutive u32 (two uint4).
__global__ void __launch_bounds__(1024) k_ldg(const uint4 *__restrict__ idx, const double *__restrict__ x, double *out, size_t iters) {
    size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    int lane = threadIdx.x & 31;
    double s = 0;
    for (size_t it = gw; it < iters; it += nw) {
        uint4 a = idx[(it * 32 + lane) * 2], b = idx[(it * 32 + lane) * 2 + 1];
        s += __ldg(x + a.x) + __ldg(x + a.y) + __ldg(x + a.z) + __ldg(x + a.w) + __ldg(x + b.x) + __ldg(x + b.y) + __ldg(x + b.z) + __ldg(x + b.w);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// MODE 0: two gather4 per lane (all 8 elements through TMA)   MODE 1: eight 16-byte bulk copies
// MODE 2: one gather4 (first 4) + four LDG (last 4)           MODE 3: four bulk copies + four LDG
template <int MODE>
__global__ void __launch_bounds__(1024) k_tma(const uint4 *__restrict__ idx, const double *__restrict__ x, const __grid_constant__ CUtensorMap map,
                                              double *out, size_t iters, int stages, int dst_stride) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int per_lane = MODE >= 2 ? dst_stride : 2 * dst_stride; // bytes of landing zone per lane and stage
    uint64_t *bars = (uint64_t *)smem;                              // nwarp * stages barriers
    unsigned char *zone = smem + 1024 + (size_t)warp * stages * 32 * per_lane;
    if (lane == 0) for (int s = 0; s < stages; s++) bar_init(&bars[warp * stages + s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    size_t gw = (size_t)blockIdx.x * nwarp + warp, nw = (size_t)gridDim.x * nwarp;
    const uint32_t tx = MODE >= 2 ? 32 * 64 : 32 * 128;
    double s = 0;
    // software pipeline: issue iteration k + stages - 1 while consuming k
    size_t it_issue = gw;
    int k_issue = 0;
    uint64_t par = 0; // 8 odd/even bits per in-flight slot (which half of the 16-byte granule is wanted)
    auto issue = [&](size_t it, int slot) {
        uint4 a = idx[(it * 32 + lane) * 2], b = make_uint4(0, 0, 0, 0);
        if (MODE < 2) b = idx[(it * 32 + lane) * 2 + 1];
        uint64_t bits = (a.x & 1) | (a.y & 1) << 1 | (a.z & 1) << 2 | (a.w & 1) << 3 | (b.x & 1) << 4 | (b.y & 1) << 5 | (b.z & 1) << 6 | (b.w & 1) << 7;
        par = (par & ~(0xFFull << (8 * slot))) | bits << (8 * slot);
        uint64_t *bar = &bars[warp * stages + slot];
        unsigned char *dst = zone + ((size_t)slot * 32 + lane) * per_lane;
        if (lane == 0) bar_expect(bar, tx);
        __syncwarp();
        if (MODE == 0) {
            tma_gather4(dst, &map, a.x >> 1, a.y >> 1, a.z >> 1, a.w >> 1, bar);
            tma_gather4(dst + dst_stride, &map, b.x >> 1, b.y >> 1, b.z >> 1, b.w >> 1, bar);
        } else if (MODE == 1) {
            bulk16(dst, x + (a.x & ~1u), bar); bulk16(dst + 16, x + (a.y & ~1u), bar); bulk16(dst + 32, x + (a.z & ~1u), bar); bulk16(dst + 48, x + (a.w & ~1u), bar);
            bulk16(dst + dst_stride, x + (b.x & ~1u), bar); bulk16(dst + dst_stride + 16, x + (b.y & ~1u), bar);
            bulk16(dst + dst_stride + 32, x + (b.z & ~1u), bar); bulk16(dst + dst_stride + 48, x + (b.w & ~1u), bar);
        } else if (MODE == 2) {
            tma_gather4(dst, &map, a.x >> 1, a.y >> 1, a.z >> 1, a.w >> 1, bar);
        } else {
            bulk16(dst, x + (a.x & ~1u), bar); bulk16(dst + 16, x + (a.y & ~1u), bar); bulk16(dst + 32, x + (a.z & ~1u), bar); bulk16(dst + 48, x + (a.w & ~1u), bar);
        }
    };
    for (; k_issue < stages - 1 && it_issue < iters; k_issue++, it_issue += nw) issue(it_issue, k_issue % stages);
    int k = 0;
    for (size_t it = gw; it < iters; it += nw, k++) {
        if (it_issue < iters) { issue(it_issue, k_issue % stages); k_issue++; it_issue += nw; }
        int slot = k % stages;
        uint32_t bits = (uint32_t)(par >> (8 * slot)) & 0xFF;
        double l0 = 0, l1 = 0, l2 = 0, l3 = 0;
        if (MODE >= 2) {
            uint4 b = idx[(it * 32 + lane) * 2 + 1];
            l0 = __ldg(x + b.x); l1 = __ldg(x + b.y); l2 = __ldg(x + b.z); l3 = __ldg(x + b.w);
        }
        bar_wait(&bars[warp * stages + slot], (k / stages) & 1);
        const double *d = (const double *)(zone + ((size_t)slot * 32 + lane) * per_lane);
        s += d[0 + (bits & 1)] + d[2 + (bits >> 1 & 1)] + d[4 + (bits >> 2 & 1)] + d[6 + (bits >> 3 & 1)];
        if (MODE >= 2) s += l0 + l1 + l2 + l3;
        else { const double *e = (const double *)((const unsigned char *)d + dst_stride); s += e[0 + (bits >> 4 & 1)] + e[2 + (bits >> 5 & 1)] + e[4 + (bits >> 6 & 1)] + e[6 + (bits >> 7 & 1)]; }
        __syncwarp();
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const char *variant = argc > 1 ? argv[1] : "ldg";
    uint32_t m = argc > 2 ? (uint32_t)atol(argv[2]) : 2000000u;
    int box_rows = argc > 3 ? atoi(argv[3]) : 1, warps = argc > 4 ? atoi(argv[4]) : 16, stages = argc > 5 ? atoi(argv[5]) : 4, dst_stride = argc > 6 ? atoi(argv[6]) : 64;
    if (stages > 8) stages = 8;
    const size_t n = (size_t)1 << 27;
    const size_t iters = n / 256;
    std::vector<uint32_t> h(n);
    uint64_t st = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i] = (uint32_t)(st % m); }
    std::vector<double> hx(m);
    for (uint32_t i = 0; i < m; i++) hx[i] = (double)(i % 1021) + 0.5;
    double want = 0; for (size_t i = 0; i < iters * 256; i++) want += hx[h[i]];
    uint32_t *idx; double *x, *out;
    const int grid = 148, tpb = warps * 32;
    CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&x, (size_t)m * 8)); CK(cudaMalloc(&out, (size_t)grid * tpb * 8));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(x, hx.data(), (size_t)m * 8, cudaMemcpyHostToDevice));
    CUtensorMap map; memset(&map, 0, sizeof map);
    int mode = !strcmp(variant, "g4") ? 0 : !strcmp(variant, "bulk") ? 1 : !strcmp(variant, "mix") ? 2 : !strcmp(variant, "mixb") ? 3 : -1;
    if (mode == 0 || mode == 2) {
        void *fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
        cuuint64_t gdim[2] = {2, (cuuint64_t)m / 2}, gstride[1] = {16};
        cuuint32_t box[2] = {2, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, x, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box rows %d)\n", (int)r, box_rows); return 4; }
    }
    size_t smem = 1024 + (size_t)warps * stages * 32 * (mode >= 2 ? dst_stride : 2 * dst_stride);
    if (mode >= 0 && smem > 227 * 1024) { printf("smem %zu too large\n", smem); return 5; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&] {
        if (mode < 0) k_ldg<<<grid, tpb>>>((const uint4 *)idx, x, out, iters);
        else if (mode == 0) k_tma<0><<<grid, tpb, smem>>>((const uint4 *)idx, x, map, out, iters, stages, dst_stride);
        else if (mode == 1) k_tma<1><<<grid, tpb, smem>>>((const uint4 *)idx, x, map, out, iters, stages, dst_stride);
        else if (mode == 2) k_tma<2><<<grid, tpb, smem>>>((const uint4 *)idx, x, map, out, iters, stages, dst_stride);
        else k_tma<3><<<grid, tpb, smem>>>((const uint4 *)idx, x, map, out, iters, stages, dst_stride);
    };
    CK(cudaFuncSetAttribute(k_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    launch(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    std::vector<double> ho((size_t)grid * tpb);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
    double got = 0; for (double v : ho) got += v;
    cudaEventRecord(e0); for (int r = 0; r < 5; r++) launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double g = iters * 256.0 / ms / 1e6;
    printf("%-5s m=%9u box_rows=%d warps=%2d stages=%d stride=%3d : %.3f ms  %.1f G gathers/s  %.2f /clk/SM @1.965GHz   sum %s (%.1f vs %.1f)\n", variant, m,
           box_rows, warps, stages, dst_stride, ms, g, g / 148 / 1.965, got == want ? "ok" : "MISMATCH", got, want);
    return got == want ? 0 : 1;
}
