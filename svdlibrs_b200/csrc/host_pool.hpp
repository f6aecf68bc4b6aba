// host_pool.hpp -- DMA-able host buffers for the SvdRec arrays (ut is d x rows(A) f64: 0.9 GB at cfg 2, 16 GB at
// cfg 4/5).  D2H into pageable memory is staged by the driver and pays first-touch page faults; into registered
// memory it runs at PCIe speed and overlaps with the kernels still computing later rows.  Blocks are anonymous
// huge-page mappings touched in parallel and then registered with CUDA (an order of magnitude faster than
// cudaHostAlloc, see host_pool.cpp); freed result buffers are kept for the next call of the process, up to
// SVDLAS2_HOST_CACHE_MB (default: the smaller of 32 GiB and 1/8 of the machine's RAM).
#pragma once
#include <cstddef>

namespace las2 {

// Returns a buffer of at least `bytes` (pinned when possible, plain malloc otherwise). Never returns null
// (throws Error(SVDLAS2_ERR_ALLOC)).
void *host_acquire(size_t bytes);
// Gives a buffer obtained from host_acquire back (cached or freed). Pointers not from the pool are free()d.
void host_release(void *p);
// Largest result (bytes) worth allocating ahead of time on a helper thread: 1/4 of the machine's RAM.
size_t host_prefetch_limit();
// Memory node the pool steers new blocks to: the node of the current CUDA device (-1: unknown / switched off).
int host_pool_node();
// Frees every cached buffer (tests; process exit does it implicitly).
void host_pool_trim();

} // namespace las2
