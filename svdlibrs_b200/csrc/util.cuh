// util.cuh -- error plumbing, RAII device buffers and small device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/svdlas2.h"

namespace las2 {

using u32 = uint32_t;
using u64 = uint64_t;

// Every failure inside the library travels as this exception up to the extern "C" boundary, which turns it
// into a status code + thread-local message (never across the ABI; reference error sites: src/error.rs:8-26).
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define LAS2_CUDA(expr)                                                                                  \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char b__[512];                                                                               \
            snprintf(b__, sizeof b__, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__,     \
                     __LINE__, cudaGetErrorString(e__));                                                 \
            throw ::las2::Error(SVDLAS2_ERR_CUDA, b__);                                                  \
        }                                                                                                \
    } while (0)

#define LAS2_LAUNCH_CHECK() LAS2_CUDA(cudaGetLastError())

// Device memory comes from a private stream-ordered pool (one per device) with the release threshold lifted, so that the buffers a
// solve frees (Lanczos panel, U/V, upload scratch) are handed back to the next allocation in microseconds instead
// of being unmapped and mapped again by cudaFree / cudaMalloc (tens of milliseconds of jitter per solve on B200).
// device_alloc returns memory that is immediately usable on any stream; device_free waits for the device like
// cudaFree does.  (util.cu)
void *device_alloc(size_t bytes);
void device_free(void *p);
void device_pool_trim(); // give cached device memory back to the driver
// host time spent inside device_alloc / device_free since the last reset: {alloc total s, alloc max s, alloc calls, free total s,
// free max s, free calls} (SVDLAS2_TRACE)
void device_alloc_stats(double out[6], bool reset);
// {number of parked blocks, their bytes} on the current device
void device_parked_stats(double out[2]);

template <typename T>
class DevBuf {
  public:
    DevBuf() = default;
    explicit DevBuf(size_t n) { alloc(n); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        n_ = n;
        if (n == 0) return;
        p_ = static_cast<T *>(device_alloc(n * sizeof(T)));
    }
    void release() {
        if (p_) device_free(p_);
        p_ = nullptr; n_ = 0;
    }
    T *get() const { return p_; }
    size_t size() const { return n_; }
    operator T *() const { return p_; }

  private:
    T *p_ = nullptr;
    size_t n_ = 0;
};

// Pinned host scalars the device writes and the host reads after a stream sync.
template <typename T>
class PinnedBuf {
  public:
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { if (p_) cudaFreeHost(p_); }
    void alloc(size_t n) {
        if (p_) cudaFreeHost(p_);
        p_ = nullptr; n_ = n;
        if (n == 0) return;
        LAS2_CUDA(cudaHostAlloc(&p_, n * sizeof(T), cudaHostAllocDefault));
    }
    T *get() const { return p_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return p_[i]; }

  private:
    T *p_ = nullptr;
    size_t n_ = 0;
};

// Function attributes (dynamic shared-memory limit, carve-out) are per device: caches of "already set" are indexed by
// the current device so that a process driving several GPUs configures every one of them.
constexpr int kMaxDevicesTracked = 64;
static inline int current_device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevicesTracked) ? dev : 0;
}

static inline unsigned cdiv(u64 a, u64 b) { return (unsigned)((a + b - 1) / b); }

constexpr int kNumSMs = 148; // B200: 2 dies x 74 SMs; grids are sized in multiples of this

#ifdef __CUDACC__
// L2 eviction policies (descriptor form: works for every load width on sm_100a; the immediate
// .L2::evict_* qualifiers are accepted by ptxas only on 256-bit loads).
__device__ __forceinline__ u64 l2_policy_evict_first() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Streaming loads: matrix values / indices are read exactly once per SpMV, so they must not displace the
// gathered vector from L1 (no_allocate) and should leave L2 first (evict_first policy).
__device__ __forceinline__ double2 ld_stream_f64x2(const double *p, u64 pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u32x2(const u32 *p, u64 pol) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;"
                 : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const u32 *p, u64 pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ u32 ld_stream_u16(const uint16_t *p, u64 pol) {
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(p), "l"(pol));
    return (u32)r;
}
__device__ __forceinline__ u32 ld_stream_u32(const u32 *p, u64 pol) {
    u32 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p, u64 pol) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
// Gathered vector entries: keep them in L1 and make them the last thing L2 evicts.
__device__ __forceinline__ double ld_gather_f64(const double *p, u64 pol) {
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

} // namespace las2
