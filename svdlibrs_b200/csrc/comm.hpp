// comm.hpp -- the one exchange step of the sharded operator: y = sum_g B_g'(B_g x), an
// all-reduce of N doubles per svd_opb over NVLink.  One process per GPU; NCCL is loaded with dlopen so the
// single-GPU library has no link-time dependency on it.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace las2 {

struct PeerRing;

class Comm {
  public:
    // id: SVDLAS2_COMM_ID_BYTES opaque bytes produced by unique_id() on rank 0 and broadcast by the caller
    static Comm *create(int rank, int world, const uint8_t *id);
    static void unique_id(uint8_t *id_out);
    ~Comm();
    int rank() const { return rank_; }
    int world() const { return world_; }
    // in-place sum over ranks, bit-identical result on every rank
    void allreduce_sum(double *buf, uint64_t n, cudaStream_t s);
    // in-place all-gather of row ranges: rank g owns buf[offs[g] .. offs[g+1]) (offs has world()+1 entries, in doubles)
    // and every rank ends up with all of them (one ncclAllGather through an internal staging buffer)
    void allgather_ranges(double *buf, const uint64_t *offs, cudaStream_t s) { allgather_ranges2(buf, nullptr, offs, s); }
    // the same for a PAIR of vectors with the same row ranges (the (q, r) pair of a row-split reorthogonalisation sweep) in
    // ONE collective: pack kernel -> ncclAllGather -> unpack kernel (b may be null)
    void allgather_ranges2(double *a, double *b, const uint64_t *offs, cudaStream_t s) {
        double *v[2] = {a, b};
        allgather_ranges_n(v, b ? 2 : 1, offs, s);
    }
    // up to 8 vectors with the same row ranges per collective (the Ritz vectors of a row-split contraction)
    void allgather_ranges_n(double *const *vecs, int nvec, const uint64_t *offs, cudaStream_t s);
    // host-side exchange of a few bytes per rank (set-up of the peer-memory ring): all[h * bytes_each ..] = rank h's block
    void allgather_host_bytes(const void *mine, void *all, size_t bytes_each);
    // all ranks have reached this point (NCCL; blocks the host)
    void barrier();
    // the peer-memory ring for vectors of n doubles (collective: every rank calls it with the same n).  One ring per
    // communicator, grown when a longer vector comes along; nullptr when peer memory is unavailable or SVDLAS2_RING=0.
    PeerRing *vector_ring(uint64_t n_doubles);

  private:
    Comm() = default;
    int rank_ = 0, world_ = 1;
    void *nccl_comm_ = nullptr;
    // short sums (reorthogonalisation coefficients, squared norms, the shared seed) go through a peer-memory ring that adds
    // in rank order (peer_ring.cuh) when peer memory is available: no floating-point result then depends on NCCL's choices
    PeerRing *small_ = nullptr;
    PeerRing *big_ = nullptr;   // the operators' vector ring (vector_ring())
    bool ring_unavailable_ = false;
    void *stage_ = nullptr; // all-gather staging (world * largest range doubles)
    uint64_t stage_cap_ = 0;
};

} // namespace las2
