// host_pool.cpp -- see host_pool.hpp.
//
// How a block is made (scripts/micro/pinned_alloc.cu, B200 box, 8 GB): cudaHostAlloc 3.5 s (2.3 GB/s, one thread
// faulting 4 KB pages inside the driver); anonymous mmap + MADV_HUGEPAGE + first touch on 16 threads 0.13 s, then
// cudaHostRegister 0.15 s (2 MB pages: 512x fewer to pin) = 28 GB/s; both copy D2H at the same 48 GB/s.
#include "host_pool.hpp"

#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "util.cuh"

namespace las2 {

namespace {

enum class Kind { Registered, Pageable };

struct Block {
    size_t cap; // bytes mapped
    Kind kind;
};

constexpr size_t kHuge = (size_t)2 << 20;

size_t phys_bytes() {
    const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
    return (pages > 0 && psz > 0) ? (size_t)pages * (size_t)psz : (size_t)64 << 30;
}

struct Pool {
    std::mutex mu;
    std::map<void *, Block> live;       // handed out
    std::multimap<size_t, void *> idle; // cached registered blocks by capacity
    size_t idle_bytes = 0;
    size_t limit() {
        static size_t lim = [] {
            const char *e = std::getenv("SVDLAS2_HOST_CACHE_MB");
            if (e) return (size_t)std::strtoull(e, nullptr, 10) << 20;
            return std::min<size_t>((size_t)32 << 30, phys_bytes() / 8);
        }();
        return lim;
    }
};

Pool &pool() {
    static Pool *p = new Pool(); // leaked on purpose: no destructor-order issues with the CUDA runtime at exit
    return *p;
}

// CPUs this process may actually burn: the cgroup quota when there is one (a burst of 16 touch threads inside a
// container with a smaller CFS quota gets the whole process throttled for the rest of the period -- seen on the
// B200 box as 0.2-0.8 s stalls of the launching thread right after an allocation), else the visible cores.
unsigned usable_cpus() {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    double quota = 0.0;
    if (FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) { // cgroup v2: "<quota|max> <period>"
        char q[64]; long period = 0;
        if (std::fscanf(f, "%63s %ld", q, &period) == 2 && period > 0 && std::strcmp(q, "max") != 0) quota = std::atof(q) / (double)period;
        std::fclose(f);
    } else if (FILE *g = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { // cgroup v1
        long q = -1, period = 100000;
        if (std::fscanf(g, "%ld", &q) != 1) q = -1;
        std::fclose(g);
        if (FILE *h = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (std::fscanf(h, "%ld", &period) != 1) period = 100000; std::fclose(h); }
        if (q > 0 && period > 0) quota = (double)q / (double)period;
    }
    if (quota >= 1.0) hw = std::min<unsigned>(hw, (unsigned)quota);
    return hw;
}

// ---- NUMA placement -------------------------------------------------------------------------------------------------
// The result blocks are the target of tens of GB of D2H traffic; on a two-socket box a block that lands on the socket the
// GPU is NOT attached to is written across the inter-socket link (8 ranks downloading 19.2 GB at cfg 4: half of them took
// 0.21 s instead of 0.13 s, profiles/r2_configs.md).  So the block is placed on the memory node of the current CUDA device:
// mbind(MPOL_PREFERRED) where the container allows it, and the first-touch threads run on that node's CPUs either way
// (first touch allocates locally).  Every step is best effort: unknown topology or a refused call leaves the old behaviour.
// SVDLAS2_HOST_NUMA=0 switches it off.
struct NodeInfo {
    int node = -1;
    cpu_set_t cpus;     // CPUs of the node that this process may run on
    bool have_cpus = false;
};

bool read_line(const std::string &path, char *buf, size_t n) {
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    const bool ok = std::fgets(buf, (int)n, f) != nullptr;
    std::fclose(f);
    return ok;
}

NodeInfo current_device_node() {
    NodeInfo ni;
    CPU_ZERO(&ni.cpus);
    if (const char *e = std::getenv("SVDLAS2_HOST_NUMA")) if (e[0] == '0') return ni;
    int dev = 0;
    char bus[64] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof bus - 1, dev) != cudaSuccess) {
        cudaGetLastError();
        return ni;
    }
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char line[4096];
    if (!read_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", line, sizeof line)) return ni;
    ni.node = std::atoi(line);
    if (ni.node < 0) return ni;
    // "0-31,64-95" -> set, intersected with what the process is allowed to use
    if (!read_line("/sys/devices/system/node/node" + std::to_string(ni.node) + "/cpulist", line, sizeof line)) return ni;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return ni;
    int found = 0;
    for (char *q = line; *q;) {
        char *end = nullptr;
        long a = std::strtol(q, &end, 10);
        if (end == q) break;
        long b = a;
        if (*end == '-') { q = end + 1; b = std::strtol(q, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++)
            if (c >= 0 && CPU_ISSET((int)c, &allowed)) { CPU_SET((int)c, &ni.cpus); found++; }
        q = (*end == ',') ? end + 1 : end;
        if (*end != ',' ) break;
    }
    ni.have_cpus = found > 0;
    return ni;
}

// returns 0 when the kernel accepted the policy
long prefer_node(void *p, size_t cap, int node) {
    if (node < 0 || node >= 1024) return -1;
    unsigned long mask[1024 / (8 * sizeof(unsigned long))] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_mbind
    return syscall(SYS_mbind, p, cap, 1 /* MPOL_PREFERRED */, mask, 1024ul + 1, 0u);
#else
    (void)p; (void)cap;
    return -1;
#endif
}

void first_touch(char *p, size_t n, bool single, const NodeInfo &ni) {
    static const unsigned want = [] {
        if (const char *e = std::getenv("SVDLAS2_TOUCH_THREADS")) return (unsigned)std::max(1, std::atoi(e));
        return std::max(1u, std::min(8u, usable_cpus() / 2)); // leave cores to the launching thread and the driver
    }();
    unsigned nt = want;
    if (n < (size_t)256 << 20 || single) nt = 1;
    auto work = [p, n, nt, &ni](unsigned t, bool own_thread) {
        // a helper thread moves to the device's memory node before it touches (the caller's affinity is left alone)
        if (own_thread && ni.have_cpus) sched_setaffinity(0, sizeof ni.cpus, &ni.cpus);
        size_t a = n / nt * t, b = (t == nt - 1) ? n : n / nt * (t + 1);
        a = a / 4096 * 4096;
        for (size_t i = a; i < b; i += 4096) p[i] = 0;
    };
    if (nt == 1 && !ni.have_cpus) { work(0, false); return; }
    std::vector<std::thread> th;
    for (unsigned t = ni.have_cpus ? 0 : 1; t < nt; t++) th.emplace_back(work, t, true);
    if (!ni.have_cpus) work(0, false);
    for (auto &x : th) x.join();
}

void unmap_block(void *p, const Block &b) {
    if (b.kind == Kind::Registered) cudaHostUnregister(p);
    munmap(p, b.cap);
}

} // namespace

size_t host_prefetch_limit() { return phys_bytes() / 4; }

int host_pool_node() { return current_device_node().node; }

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
            P.live[p] = Block{cap, Kind::Registered};
            return p;
        }
    }
    const size_t cap = (bytes + kHuge - 1) / kHuge * kHuge;
    // debugging switches (SVDLAS2_HOST_POOL): "nothp" = no huge-page advice, "touch1" = single-threaded first touch
    static const std::string mode = [] { const char *e = std::getenv("SVDLAS2_HOST_POOL"); return std::string(e ? e : ""); }();
    void *p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED)
        throw Error(SVDLAS2_ERR_ALLOC, "host result allocation of " + std::to_string(bytes) + " bytes failed");
    if (mode != "nothp") madvise(p, cap, MADV_HUGEPAGE); // advisory: without THP the block is still correct, only slower to pin
    const bool trace = std::getenv("SVDLAS2_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const NodeInfo ni = current_device_node();
    const long bound = ni.node >= 0 ? prefer_node(p, cap, ni.node) : -1;
    first_touch((char *)p, cap, mode == "touch1", ni);
    const auto t1 = std::chrono::steady_clock::now();
    Kind kind = Kind::Registered;
    if (cudaHostRegister(p, cap, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        kind = Kind::Pageable; // D2H still works (staged by the driver), just slower
    }
    if (trace)
        fprintf(stderr, "[svdlas2] host pool: new block of %.3f GB: first touch %.3f s, cudaHostRegister %.3f s (%s); memory node of the "
                        "device %d (mbind %s, touch threads %s)\n",
                (double)cap / 1e9, std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(),
                kind == Kind::Registered ? "registered" : "pageable", ni.node, ni.node < 0 ? "skipped" : bound == 0 ? "ok" : "refused",
                ni.have_cpus ? "on that node" : "unbound");
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = Block{cap, kind};
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
        if (b.kind == Kind::Registered && b.cap <= P.limit()) {
            // make room (evict smallest first), then cache
            while (P.idle_bytes + b.cap > P.limit() && !P.idle.empty()) {
                auto v = P.idle.begin();
                P.idle_bytes -= v->first;
                unmap_block(v->second, Block{v->first, Kind::Registered});
                P.idle.erase(v);
            }
            if (P.idle_bytes + b.cap <= P.limit()) {
                P.idle.emplace(b.cap, p);
                P.idle_bytes += b.cap;
                return;
            }
        }
    }
    unmap_block(p, b);
}

void host_pool_trim() {
    Pool &P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    for (auto &kv : P.idle) unmap_block(kv.second, Block{kv.first, Kind::Registered});
    P.idle.clear();
    P.idle_bytes = 0;
}

} // namespace las2
