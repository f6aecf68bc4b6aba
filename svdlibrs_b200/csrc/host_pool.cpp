// host_pool.cpp -- see host_pool.hpp.
#include "host_pool.hpp"

#include <cstdlib>
#include <map>
#include <mutex>

#include "util.cuh"

namespace las2 {

namespace {

struct Block {
    size_t cap;
    bool pinned;
};

struct Pool {
    std::mutex mu;
    std::map<void *, Block> live;            // handed out
    std::multimap<size_t, void *> idle;      // cached pinned blocks by capacity
    size_t idle_bytes = 0;
    size_t limit() {
        static size_t lim = [] {
            const char *e = std::getenv("SVDLAS2_HOST_CACHE_MB");
            return (size_t)(e ? std::strtoull(e, nullptr, 10) : 4096) << 20;
        }();
        return lim;
    }
};

Pool &pool() {
    static Pool *p = new Pool(); // leaked on purpose: no destructor-order issues with the CUDA runtime at exit
    return *p;
}

} // namespace

void *host_acquire(size_t bytes) {
    if (bytes == 0) bytes = 8;
    Pool &P = pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.idle.lower_bound(bytes);
        if (it != P.idle.end() && it->first <= bytes + bytes / 2 + (1 << 20)) {
            void *p = it->second;
            size_t cap = it->first;
            P.idle.erase(it);
            P.idle_bytes -= cap;
            P.live[p] = Block{cap, true};
            return p;
        }
    }
    void *p = nullptr;
    bool pinned = true;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        pinned = false;
        p = std::malloc(bytes);
        if (!p) throw Error(SVDLAS2_ERR_ALLOC, "host result allocation of " + std::to_string(bytes) + " bytes failed");
    }
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = Block{bytes, pinned};
    return p;
}

void host_release(void *p) {
    if (!p) return;
    Pool &P = pool();
    Block b;
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.live.find(p);
        if (it == P.live.end()) { std::free(p); return; }
        b = it->second;
        P.live.erase(it);
        if (b.pinned && b.cap <= P.limit()) {
            // make room (evict smallest first), then cache
            while (P.idle_bytes + b.cap > P.limit() && !P.idle.empty()) {
                auto v = P.idle.begin();
                P.idle_bytes -= v->first;
                cudaFreeHost(v->second);
                P.idle.erase(v);
            }
            if (P.idle_bytes + b.cap <= P.limit()) {
                P.idle.emplace(b.cap, p);
                P.idle_bytes += b.cap;
                return;
            }
        }
    }
    if (b.pinned) cudaFreeHost(p); else std::free(p);
}

void host_pool_trim() {
    Pool &P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    for (auto &kv : P.idle) cudaFreeHost(kv.second);
    P.idle.clear();
    P.idle_bytes = 0;
}

} // namespace las2
