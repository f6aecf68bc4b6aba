// Microbenchmark: how fast can a large HOST result buffer be made DMA-able?
//   (a) cudaHostAlloc, (b) malloc + parallel first touch + cudaHostRegister, (c) mmap + MADV_HUGEPAGE + parallel touch +
//   cudaHostRegister, (d) plain malloc + parallel touch (no pinning) followed by a pageable cudaMemcpy D2H.
// nvcc -O3 -o pinned_alloc pinned_alloc.cu -lpthread && ./pinned_alloc [GB]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void touch(char *p, size_t n, int nt) {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([=] { size_t a = n / nt * t, b = (t == nt - 1) ? n : n / nt * (t + 1); for (size_t i = a; i < b; i += 4096) p[i] = 0; });
    for (auto &x : th) x.join();
}
int main(int argc, char **argv) {
    const size_t gb = argc > 1 ? atoi(argv[1]) : 4, n = gb << 30;
    const int nt = std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    cudaFree(0);
    void *d; cudaMalloc(&d, n); cudaMemset(d, 1, n); cudaDeviceSynchronize();
    auto d2h = [&](void *h) { double t = now(); cudaMemcpy(h, d, n, cudaMemcpyDeviceToHost); return now() - t; };
    { double t = now(); void *p; cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocPortable); double dt = now() - t;
      printf("(a) cudaHostAlloc %zu GB: %.3f s (%.2f GB/s) %s; D2H %.3f s\n", gb, dt, gb / dt, cudaGetErrorString(e), d2h(p)); t = now(); cudaFreeHost(p); printf("    cudaFreeHost %.3f s\n", now() - t); }
    { double t = now(); char *p = (char *)malloc(n); touch(p, n, nt); double t1 = now() - t; t = now(); cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterPortable); double dt = now() - t;
      printf("(b) malloc + touch(%d thr) %.3f s, cudaHostRegister %.3f s %s; D2H %.3f s\n", nt, t1, dt, cudaGetErrorString(e), d2h(p)); t = now(); cudaHostUnregister(p); free(p); printf("    unregister+free %.3f s\n", now() - t); }
    { double t = now(); char *p = (char *)mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0); madvise(p, n, MADV_HUGEPAGE); touch(p, n, nt); double t1 = now() - t; t = now();
      cudaError_t e = cudaHostRegister(p, n, cudaHostRegisterPortable); double dt = now() - t;
      printf("(c) mmap + MADV_HUGEPAGE + touch(%d thr) %.3f s, cudaHostRegister %.3f s %s; D2H %.3f s\n", nt, t1, dt, cudaGetErrorString(e), d2h(p)); t = now(); cudaHostUnregister(p); munmap(p, n); printf("    unregister+munmap %.3f s\n", now() - t); }
    { double t = now(); char *p = (char *)malloc(n); touch(p, n, nt); double t1 = now() - t;
      printf("(d) malloc + touch %.3f s; pageable D2H %.3f s\n", t1, d2h(p)); free(p); }
    return 0;
}
