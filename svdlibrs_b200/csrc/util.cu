// util.cu -- device memory pool behind DevBuf (see util.cuh).
#include "util.cuh"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace las2 {

namespace {

constexpr int kMaxDevices = 64;
struct PoolState {
    std::mutex mu;
    cudaStream_t stream[kMaxDevices] = {};
    cudaMemPool_t pool[kMaxDevices] = {};
    bool ready[kMaxDevices] = {};
    bool legacy = false; // SVDLAS2_DEVICE_POOL=0: plain cudaMalloc / cudaFree
    // Large blocks (>= kBigBlock) are not handed back to the stream-ordered pool but parked here and re-issued to the next
    // request of (nearly) the same size.  Why: the pool keeps its memory (release threshold lifted), yet every few solves a
    // request of the same 0.8 GB that was freed a moment ago came back after 0.1-0.43 s -- the pool had carved smaller
    // allocations out of the freed range and mapped fresh physical memory for the big one (40 solves of cfg 2: six of them
    // 2-5 x slower for this one allocation alone, profiles/r2ab_allocator_outliers.md).  A solve repeats its sequence of
    // sizes exactly, so an exact-fit list makes the reuse deterministic.  svdlas2_trim_caches() and an allocation failure
    // release everything that is parked.
    struct Parked { void *p; size_t cap; uint64_t tick; };
    std::vector<Parked> parked[kMaxDevices];
    size_t parked_bytes[kMaxDevices] = {};
    std::unordered_map<void *, size_t> big_live; // blocks >= kBigBlock handed out: pointer -> capacity
    // bytes worth keeping parked: unlimited by default -- the pool these blocks would otherwise return to keeps its memory
    // as well (release threshold lifted), so parking adds no footprint; SVDLAS2_DEVICE_PARK_MB caps it.  A cap makes a cycle
    // of requests larger than it thrash (oldest out first), which is what the e2e call of cfg 4 did with a third of the device.
    size_t park_limit[kMaxDevices] = {};
    uint64_t tick = 0;                           // big requests so far: parked blocks nobody asked for in kParkAge requests go back
};
constexpr uint64_t kParkAge = 4096;
constexpr size_t kBigBlock = (size_t)8 << 20;

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

// SVDLAS2_TRACE bookkeeping: where a host thread loses time inside the allocator (device_alloc_stats)
std::atomic<uint64_t> g_alloc_ns{0}, g_alloc_max_ns{0}, g_alloc_calls{0}, g_free_ns{0}, g_free_max_ns{0}, g_free_calls{0};
struct AllocTimer {
    std::atomic<uint64_t> &sum, &mx, &calls;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~AllocTimer() {
        const uint64_t ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        sum += ns; calls += 1;
        uint64_t cur = mx.load();
        while (ns > cur && !mx.compare_exchange_weak(cur, ns)) {}
    }
};

} // namespace

void device_parked_stats(double out[2]) {
    int dev = 0;
    out[0] = out[1] = 0.0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= kMaxDevices) return;
    PoolState &S = state();
    std::lock_guard<std::mutex> lk(S.mu);
    out[0] = (double)S.parked[dev].size();
    out[1] = (double)S.parked_bytes[dev];
}

void device_alloc_stats(double out[6], bool reset) {
    out[0] = g_alloc_ns.load() * 1e-9; out[1] = g_alloc_max_ns.load() * 1e-9; out[2] = (double)g_alloc_calls.load();
    out[3] = g_free_ns.load() * 1e-9; out[4] = g_free_max_ns.load() * 1e-9; out[5] = (double)g_free_calls.load();
    if (reset) { g_alloc_ns = 0; g_alloc_max_ns = 0; g_alloc_calls = 0; g_free_ns = 0; g_free_max_ns = 0; g_free_calls = 0; }
}

namespace {

void *pool_alloc(int dev, size_t bytes, cudaError_t *err) {
    void *p = nullptr;
    cudaStream_t as = alloc_stream(dev);
    cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, state().pool[dev], as);
    if (e == cudaSuccess) e = cudaStreamSynchronize(as);
    if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; }
    *err = e;
    return p;
}

void pool_free(int dev, void *p) {
    cudaStream_t as = nullptr;
    try { as = alloc_stream(dev); } catch (...) { cudaFree(p); return; }
    if (cudaFreeAsync(p, as) != cudaSuccess) { cudaGetLastError(); cudaFree(p); }
}

// hands every parked block of the device back to the pool (the caller has made sure the device is idle)
void unpark_all(int dev) {
    PoolState &S = state();
    std::vector<PoolState::Parked> take;
    {
        std::lock_guard<std::mutex> lk(S.mu);
        take.swap(S.parked[dev]);
        S.parked_bytes[dev] = 0;
    }
    for (auto &b : take) pool_free(dev, b.p);
}

} // namespace

void device_pool_trim();

void *device_alloc(size_t bytes) {
    void *p = nullptr;
    if (bytes == 0) return nullptr;
    AllocTimer timer{g_alloc_ns, g_alloc_max_ns, g_alloc_calls};
    cudaError_t e;
    int dev = 0;
    LAS2_CUDA(cudaGetDevice(&dev));
    if (state().legacy || dev >= kMaxDevices) {
        e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) cudaGetLastError();
    } else {
        PoolState &S = state();
        if (bytes >= kBigBlock) {
            // smallest parked block that fits without wasting more than 1/8 of it
            std::lock_guard<std::mutex> lk(S.mu);
            auto &v = S.parked[dev];
            size_t best = v.size();
            for (size_t i = 0; i < v.size(); i++)
                if (v[i].cap >= bytes && v[i].cap - bytes <= bytes / 8 && (best == v.size() || v[i].cap < v[best].cap)) best = i;
            if (best != v.size()) {
                const PoolState::Parked b = v[best];
                v.erase(v.begin() + (long)best);
                S.parked_bytes[dev] -= b.cap;
                S.big_live[b.p] = b.cap;
                S.tick++;
                return b.p;
            }
            S.tick++;
        }
        p = pool_alloc(dev, bytes, &e);
        if (!p) { // out of memory with blocks parked: give them back to the driver and try once more
            device_pool_trim();
            p = pool_alloc(dev, bytes, &e);
        }
        if (p && bytes >= kBigBlock) {
            std::lock_guard<std::mutex> lk(S.mu);
            S.big_live[p] = bytes;
        }
    }
    if (!p)
        throw Error(SVDLAS2_ERR_ALLOC, "device allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    return p;
}

void device_free(void *p) {
    if (!p) return;
    AllocTimer timer{g_free_ns, g_free_max_ns, g_free_calls};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || state().legacy || dev >= kMaxDevices) { cudaFree(p); return; }
    // like cudaFree: nothing that may still touch the buffer is in flight once the device is idle
    cudaDeviceSynchronize();
    PoolState &S = state();
    std::vector<PoolState::Parked> evict; // handed back to the pool outside the lock
    {
        std::lock_guard<std::mutex> lk(S.mu);
        auto it = S.big_live.find(p);
        if (it != S.big_live.end()) {
            const size_t cap = it->second;
            S.big_live.erase(it);
            if (S.park_limit[dev] == 0) {
                const char *e = std::getenv("SVDLAS2_DEVICE_PARK_MB");
                S.park_limit[dev] = e ? (((size_t)std::strtoull(e, nullptr, 10) << 20) | 1) : ~(size_t)0;
            }
            if (cap <= S.park_limit[dev]) {
                // make room: the oldest parked blocks go back to the pool first
                auto &v = S.parked[dev];
                while (!v.empty() && (S.parked_bytes[dev] > S.park_limit[dev] - cap || S.tick - v.front().tick > kParkAge)) {
                    evict.push_back(v.front());
                    S.parked_bytes[dev] -= v.front().cap;
                    v.erase(v.begin());
                }
                v.push_back(PoolState::Parked{p, cap, S.tick});
                S.parked_bytes[dev] += cap;
                p = nullptr;
            }
        }
    }
    for (auto &b : evict) pool_free(dev, b.p);
    if (p) pool_free(dev, p);
}

void device_pool_trim() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceSynchronize();
    if (dev < kMaxDevices && state().ready[dev]) {
        unpark_all(dev);
        cudaStreamSynchronize(state().stream[dev]); // the releases above are ordered on the allocation stream
        cudaMemPoolTrimTo(state().pool[dev], 0);
    }
}

} // namespace las2
