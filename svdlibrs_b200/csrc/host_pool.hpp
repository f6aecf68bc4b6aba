// host_pool.hpp -- pinned host buffers for the SvdRec arrays (ut is d x rows(A) f64: 0.9 GB at cfg 2, 16 GB at
// cfg 4).  D2H into pageable memory runs at ~2 GB/s (staging + first-touch page faults); into pinned memory it
// runs at PCIe speed and overlaps with the kernels still computing later rows.  Pinning is expensive, so freed
// result buffers are kept (up to SVDLAS2_HOST_CACHE_MB, default 4096 MiB) for the next call of the process.
#pragma once
#include <cstddef>

namespace las2 {

// Returns a buffer of at least `bytes` (pinned when possible, plain malloc otherwise). Never returns null
// (throws Error(SVDLAS2_ERR_ALLOC)).
void *host_acquire(size_t bytes);
// Gives a buffer obtained from host_acquire back (cached or freed). Pointers not from the pool are free()d.
void host_release(void *p);
// Frees every cached buffer (tests; process exit does it implicitly).
void host_pool_trim();

} // namespace las2
