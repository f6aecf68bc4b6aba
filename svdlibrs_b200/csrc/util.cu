// util.cu -- device memory pool behind DevBuf (see util.cuh).
#include "util.cuh"

#include <cstdint>
#include <mutex>

namespace las2 {

namespace {

constexpr int kMaxDevices = 64;
struct PoolState {
    std::mutex mu;
    cudaStream_t stream[kMaxDevices] = {};
    cudaMemPool_t pool[kMaxDevices] = {};
    bool ready[kMaxDevices] = {};
    bool legacy = false; // SVDLAS2_DEVICE_POOL=0: plain cudaMalloc / cudaFree
};

PoolState &state() {
    static PoolState *s = [] {
        auto *p = new PoolState(); // leaked on purpose (no destructor-order issues with the CUDA runtime at exit)
        const char *e = std::getenv("SVDLAS2_DEVICE_POOL");
        p->legacy = e && e[0] == '0';
        return p;
    }();
    return *s;
}

// an otherwise idle stream per device: allocations ordered on it are available as soon as it has been synchronised
cudaStream_t alloc_stream(int dev) {
    PoolState &S = state();
    std::lock_guard<std::mutex> lk(S.mu);
    if (!S.ready[dev]) {
        LAS2_CUDA(cudaStreamCreateWithFlags(&S.stream[dev], cudaStreamNonBlocking));
        // a PRIVATE pool: the release threshold is lifted on it, not on the device's default pool that the host
        // application's own cudaMallocAsync calls use
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        LAS2_CUDA(cudaMemPoolCreate(&S.pool[dev], &props));
        uint64_t keep = UINT64_MAX;
        LAS2_CUDA(cudaMemPoolSetAttribute(S.pool[dev], cudaMemPoolAttrReleaseThreshold, &keep));
        S.ready[dev] = true;
    }
    return S.stream[dev];
}

} // namespace

void *device_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) return nullptr;
    cudaError_t e;
    int dev = 0;
    LAS2_CUDA(cudaGetDevice(&dev));
    if (state().legacy || dev >= kMaxDevices) {
        e = cudaMalloc(&p, bytes);
    } else {
        cudaStream_t as = alloc_stream(dev);
        e = cudaMallocFromPoolAsync(&p, bytes, state().pool[dev], as);
        if (e == cudaSuccess) e = cudaStreamSynchronize(as);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw Error(SVDLAS2_ERR_ALLOC, "device allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    return p;
}

void device_free(void *p) {
    if (!p) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || state().legacy || dev >= kMaxDevices) { cudaFree(p); return; }
    // like cudaFree: nothing that may still touch the buffer is in flight once the device is idle
    cudaDeviceSynchronize();
    cudaStream_t as = nullptr;
    try { as = alloc_stream(dev); } catch (...) { cudaFree(p); return; }
    if (cudaFreeAsync(p, as) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
}

void device_pool_trim() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceSynchronize();
    if (dev < kMaxDevices && state().ready[dev]) cudaMemPoolTrimTo(state().pool[dev], 0);
}

} // namespace las2
