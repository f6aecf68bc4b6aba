// comm.cpp -- NCCL all-reduce over NVLink / NVSwitch for the row-sharded operator (see comm.hpp).
// NCCL 2.28 is resolved at run time: first among the libraries already mapped into the process (the torch
// wheel's libnccl.so.2 when the host program is Python), then by soname.
#include "comm.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>

#include "peer_ring.cuh"
#include "util.cuh"

namespace las2 {

namespace {

struct NcclUniqueId { char internal[128]; };
using ncclComm_t = void *;
enum { kNcclSum = 0, kNcclInt8 = 0, kNcclFloat64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.why = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.AllGather && a.CommDestroy && a.GetErrorString;
        if (!a.ok) a.why = "libnccl.so.2 lacks a required symbol";
    });
    return a;
}

void check(int rc, const char *what) {
    if (rc != 0) throw Error(SVDLAS2_ERR_COMM, std::string(what) + ": " + api().GetErrorString(rc));
}

} // namespace

void Comm::unique_id(uint8_t *id_out) {
    NcclApi &a = api();
    if (!a.ok) throw Error(SVDLAS2_ERR_COMM, a.why);
    NcclUniqueId id;
    check(a.GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof id == SVDLAS2_COMM_ID_BYTES, "id size");
    std::memcpy(id_out, id.internal, sizeof id);
}

Comm *Comm::create(int rank, int world, const uint8_t *id_bytes) {
    if (world < 1 || rank < 0 || rank >= world) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "bad rank / world_size");
    Comm *c = new Comm();
    c->rank_ = rank;
    c->world_ = world;
    if (world == 1) return c;
    NcclApi &a = api();
    if (!a.ok) { delete c; throw Error(SVDLAS2_ERR_COMM, a.why); }
    NcclUniqueId id;
    std::memcpy(id.internal, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    int rc = a.CommInitRank(&comm, world, id, rank);
    if (rc != 0) { delete c; check(rc, "ncclCommInitRank"); }
    c->nccl_comm_ = comm;
    const char *ring_env = std::getenv("SVDLAS2_RING");
    if (!(ring_env && ring_env[0] == '0')) {
        try { c->small_ = PeerRing::create(c, 0); } catch (const Error &) { c->small_ = nullptr; } // collective, same verdict everywhere
    }
    return c;
}

void Comm::barrier() {
    if (world_ == 1 || !nccl_comm_) return;
    unsigned char mine = 1, all[8 * 8] = {0};
    allgather_host_bytes(&mine, all, 1);
}

PeerRing *Comm::vector_ring(uint64_t n_doubles) {
    if (world_ == 1 || ring_unavailable_) return nullptr;
    const char *ring_env = std::getenv("SVDLAS2_RING");
    if (ring_env && ring_env[0] == '0') return nullptr;
    if (big_ && big_->cap >= n_doubles) return big_;
    if (big_) { // grow: every rank un-maps the old blocks before anybody frees its own
        big_->close_peers();
        barrier();
        delete big_;
        big_ = nullptr;
    }
    try { big_ = PeerRing::create(this, n_doubles); }
    catch (const Error &e) {
        big_ = nullptr;
        ring_unavailable_ = true; // same verdict on every rank (PeerRing::create agrees on it collectively)
        if (std::getenv("SVDLAS2_TRACE")) fprintf(stderr, "[svdlas2] %s -- using ncclAllReduce\n", e.what());
    }
    return big_;
}

Comm::~Comm() {
    if (big_) {
        big_->close_peers();
        try { barrier(); } catch (...) {}
        delete big_;
    }
    if (small_) {
        // every rank un-maps the others' blocks before anybody frees its own
        small_->close_peers();
        try { barrier(); } catch (...) {}
        delete small_;
    }
    if (stage_) cudaFree(stage_);
    if (nccl_comm_) api().CommDestroy(nccl_comm_);
}

void Comm::allreduce_sum(double *buf, uint64_t n, cudaStream_t s) {
    if (world_ == 1) return;
    if (small_ && n <= kRingSmallCap) { small_->allreduce_small(buf, n, s); return; }
    check(api().AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, nccl_comm_, s), "ncclAllReduce");
}

void Comm::allgather_host_bytes(const void *mine, void *all, size_t bytes_each) {
    if (world_ == 1) { std::memcpy(all, mine, bytes_each); return; }
    void *d = nullptr;
    LAS2_CUDA(cudaMalloc(&d, bytes_each * (size_t)world_));
    cudaError_t e = cudaMemcpy(static_cast<char *>(d) + (size_t)rank_ * bytes_each, mine, bytes_each, cudaMemcpyHostToDevice);
    int rc = 0;
    if (e == cudaSuccess) {
        rc = api().AllGather(static_cast<char *>(d) + (size_t)rank_ * bytes_each, d, bytes_each, kNcclInt8, nccl_comm_, (cudaStream_t)0);
        if (rc == 0) e = cudaStreamSynchronize((cudaStream_t)0);
        if (rc == 0 && e == cudaSuccess) e = cudaMemcpy(all, d, bytes_each * (size_t)world_, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    check(rc, "ncclAllGather (host bytes)");
    LAS2_CUDA(e);
}

namespace {

struct GatherPlan { uint64_t offs[9]; double *vec[8]; int world, rank, nvec; uint64_t slot; };

// stage[(g * nvec + v) * slot + i] <-> vec_v[offs[g] + i]; pack: own range only; unpack: every foreign range
__global__ void k_stage_ranges(GatherPlan p, double *__restrict__ stage, int unpack) {
    const int g = unpack ? (int)blockIdx.y : p.rank;
    if (unpack && g == p.rank) return;
    const uint64_t cnt = p.offs[g + 1] - p.offs[g];
    for (int v = 0; v < p.nvec; v++) {
        double *vec = p.vec[v] + p.offs[g];
        double *st = stage + ((uint64_t)g * p.nvec + v) * p.slot;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (uint64_t)gridDim.x * blockDim.x) {
            if (unpack) vec[i] = st[i]; else st[i] = vec[i];
        }
    }
}

} // namespace

void Comm::allgather_ranges_n(double *const *vecs, int nvec, const uint64_t *offs, cudaStream_t s) {
    if (world_ == 1 || nvec < 1) return;
    if (world_ > 8 || nvec > 8) throw Error(SVDLAS2_ERR_COMM, "allgather_ranges: more than 8 ranks / vectors");
    NcclApi &api_ = api();
    // one ncclAllGather over equal slots of a staging buffer (eight grouped broadcasts cost milliseconds on NVSwitch;
    // one all-gather of 16 MB ~60 us); own ranges are packed and foreign ranges unpacked by one kernel each
    GatherPlan p;
    p.world = world_; p.rank = rank_; p.nvec = nvec; p.slot = 0;
    for (int v = 0; v < 8; v++) p.vec[v] = v < nvec ? vecs[v] : nullptr;
    for (int g = 0; g <= world_; g++) p.offs[g] = offs[g];
    for (int g = 0; g < world_; g++) p.slot = std::max(p.slot, offs[g + 1] - offs[g]);
    if (p.slot == 0) return;
    const uint64_t per_rank = p.slot * (uint64_t)p.nvec;
    if (stage_cap_ < per_rank * (uint64_t)world_) {
        if (stage_) { LAS2_CUDA(cudaStreamSynchronize(s)); cudaFree(stage_); }
        stage_ = nullptr;
        stage_cap_ = per_rank * (uint64_t)world_;
        LAS2_CUDA(cudaMalloc(&stage_, stage_cap_ * sizeof(double)));
    }
    double *st = static_cast<double *>(stage_);
    const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((p.slot + 1023) / 1024, 148));
    k_stage_ranges<<<dim3(gx, 1), 256, 0, s>>>(p, st, 0);
    check(api_.AllGather(st + (uint64_t)rank_ * per_rank, st, per_rank, kNcclFloat64, nccl_comm_, s), "ncclAllGather");
    k_stage_ranges<<<dim3(gx, world_), 256, 0, s>>>(p, st, 1);
    LAS2_LAUNCH_CHECK();
}

} // namespace las2
