// vecops.cu -- fused, deterministic BLAS-1 kernels (see vecops.cuh for the conventions).
// All kernels are HBM/L2-bound streams over N-vectors: 128-bit loads, grid capped at 4 CTAs per SM.
#include "vecops.cuh"

#include <algorithm>

namespace las2 {

namespace {

// Canonical sum of n per-CTA partials: warp 0, lane-strided sequential adds, then a shuffle tree.
// Every kernel that needs the scalar calls exactly this, so all of them see the same bits.
__device__ __forceinline__ double cta_sum_partials(const double *__restrict__ p, int n, double *smem1) {
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += 32) s += p[i];
        s = warp_sum(s);
        if (threadIdx.x == 0) *smem1 = s;
    }
    __syncthreads();
    return *smem1;
}

__device__ __forceinline__ double cta_reduce(double v, double *smem /* >= blockDim/32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; w++) s += smem[w];
    }
    return s; // valid in thread 0
}

__global__ void __launch_bounds__(kVecThreads) k_scale(double2 *__restrict__ dst, const double2 *__restrict__ src,
                                                       double alpha, u64 n2) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) {
        double2 v = src[i];
        v.x = __dmul_rn(alpha, v.x);
        v.y = __dmul_rn(alpha, v.y);
        dst[i] = v;
    }
}

__global__ void __launch_bounds__(kVecThreads) k_dot(const double2 *__restrict__ a, const double2 *__restrict__ b,
                                                     u64 n2, double *__restrict__ out) {
    __shared__ double sm[kVecThreads / 32];
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (; i < n2; i += stride) {
        double2 av = a[i], bv = b[i];
        acc = fma(av.x, bv.x, acc);
        acc = fma(av.y, bv.y, acc);
    }
    double s = cta_reduce(acc, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// r -= coef * x ; partial of r . y (y == nullptr -> r . r ; DOT == false -> no reduction at all)
template <bool DOT>
__global__ void __launch_bounds__(kVecThreads)
k_axpy_dot(double2 *__restrict__ r, const double2 *__restrict__ x, Coef coef, const double2 *__restrict__ y, u64 n2,
           double *__restrict__ out) {
    __shared__ double sm[kVecThreads / 32];
    __shared__ double cs;
    const double c = coef.parts ? cta_sum_partials(coef.parts, coef.nparts, &cs) : coef.imm;
    const double nc = -c;
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (; i < n2; i += stride) {
        double2 rv = r[i];
        const double2 xv = x[i];
        rv.x = __dadd_rn(rv.x, __dmul_rn(nc, xv.x)); // y += da * x with da = -t, as svd_daxpy(-t, x, y)
        rv.y = __dadd_rn(rv.y, __dmul_rn(nc, xv.y));
        r[i] = rv;
        if (DOT) {
            double2 yv;
            if (y == nullptr) yv = rv;
            else if (y == x) yv = xv;
            else yv = y[i];
            acc = fma(rv.x, yv.x, acc);
            acc = fma(rv.y, yv.y, acc);
        }
    }
    if (DOT) {
        double s = cta_reduce(acc, sm);
        if (threadIdx.x == 0) out[blockIdx.x] = s;
    }
}

// partial sums of a . a for a vector of arbitrary length and alignment (the rows of U are M doubles apart)
__global__ void __launch_bounds__(kVecThreads) k_sumsq_any(const double *__restrict__ a, u64 n, double *__restrict__ out) {
    __shared__ double sm[kVecThreads / 32];
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (; i < n; i += stride) {
        const double v = a[i];
        acc = fma(v, v, acc);
    }
    double s = cta_reduce(acc, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// sigma = sqrt(t), u *= 1/sigma  (ritvec tail, src/lib.rs:1851, 1858); t arrives as partials
__global__ void __launch_bounds__(kVecThreads)
k_sigma_scale(double *__restrict__ u, u64 m, const double *__restrict__ parts, int nparts, double *__restrict__ sigma_out) {
    __shared__ double ts;
    const double t = cta_sum_partials(parts, nparts, &ts);
    const double sigma = sqrt(t);
    const double inv = 1.0 / sigma;
    if (blockIdx.x == 0 && threadIdx.x == 0) *sigma_out = sigma;
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < m; i += stride) u[i] = __dmul_rn(u[i], inv);
}

__global__ void k_gather_scalars(GatherList g, double *__restrict__ out) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < g.count) {
        double s = 0.0;
        const double *p = g.parts[w];
        const int n = g.nparts[w];
        for (int i = lane; i < n; i += 32) s += p[i];
        s = warp_sum(s);
        if (lane == 0) out[w] = s;
    }
}

// Grid-wide barrier of a launch whose CTAs are all resident (grid <= number of SMs, one small CTA each): arrive on a counter,
// the last one resets it and bumps the generation the others spin on.  (cooperative_groups' grid.sync() cost 10-40 us per
// barrier here with 100-500 CTAs: profiles/r2_step_chain.md.)
__device__ __forceinline__ void chain_barrier(unsigned *count, unsigned *gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned my_gen = *reinterpret_cast<volatile unsigned *>(gen);
        if (atomicAdd(count, 1u) == gridDim.x - 1) {
            *reinterpret_cast<volatile unsigned *>(count) = 0u;
            __threadfence();
            atomicAdd(gen, 1u);
        } else {
            while (*reinterpret_cast<volatile unsigned *>(gen) == my_gen) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// y_row = sum_k val[k] * x[idx[k]] for the rows of a CSR matrix, one sub-warp of L lanes per row; emit(row, value) runs on the
// sub-warp's first lane.  COHERENT: x was written earlier in this launch by other CTAs (read it through L2, not the
// non-coherent path).  All 32 lanes of a warp make the same number of trips (the shuffles need them all).
template <int L, bool COHERENT, class Emit>
__device__ __forceinline__ void subwarp_spmv(const u32 *__restrict__ off, const u32 *__restrict__ idx,
                                             const double *__restrict__ val, u32 rows, const double *x, Emit emit) {
    const u32 lane = threadIdx.x & (L - 1);
    const u32 sub = ((u32)blockIdx.x * blockDim.x + threadIdx.x) / L;
    const u32 nsub = (u32)gridDim.x * blockDim.x / L;
    const u32 sub_in_warp = (threadIdx.x & 31u) / L;
    for (u32 base = sub - sub_in_warp; base < rows; base += nsub) {
        const u32 row = base + sub_in_warp;
        u32 b = 0, e = 0;
        if (row < rows) { b = off[row]; e = off[row + 1]; }
        double acc = 0.0;
        for (u32 k = b + lane; k < e; k += L) {
            const double xv = COHERENT ? __ldcg(x + idx[k]) : __ldg(x + idx[k]);
            acc = fma(val[k], xv, acc);
        }
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && row < rows) emit(row, acc);
    }
}

template <bool COHERENT, class Emit>
__device__ __forceinline__ void subwarp_spmv_any(int lanes, const u32 *off, const u32 *idx, const double *val, u32 rows,
                                                 const double *x, Emit emit) {
    switch (lanes) {
    case 2: subwarp_spmv<2, COHERENT>(off, idx, val, rows, x, emit); break;
    case 4: subwarp_spmv<4, COHERENT>(off, idx, val, rows, x, emit); break;
    case 8: subwarp_spmv<8, COHERENT>(off, idx, val, rows, x, emit); break;
    case 16: subwarp_spmv<16, COHERENT>(off, idx, val, rows, x, emit); break;
    default: subwarp_spmv<32, COHERENT>(off, idx, val, rows, x, emit); break;
    }
}

template <bool SPMV>
__global__ void __launch_bounds__(1024)
k_step_chain(double2 *__restrict__ r, u64 n2, StepChain c, StepSpmv m, double *__restrict__ work, double *__restrict__ host_out,
             unsigned long long seq) {
    unsigned *bar = reinterpret_cast<unsigned *>(work + 6 * (size_t)kMaxParts);
    __shared__ double sm[32];
    __shared__ double cs;
    double coef_seen[6];
    const double *parts = c.alf_parts;
    int nparts = c.alf_nparts;
    int k0 = 0;
    if (SPMV) {
        // ---- svd_opb and the step's first update in place of the chain's phase 0 ----
        const double *x = c.y[0];     // q_j
        const double *qprev = c.x[0]; // q_{j-1}
        double *t = m.t;
        subwarp_spmv_any<false>(m.b_lanes, m.b_off, m.b_idx, m.b_val, m.b_rows, x, [t](u32 row, double v) { t[row] = v; });
        chain_barrier(bar, bar + 1);
        const double nc = -c.coef0;
        double *rs = reinterpret_cast<double *>(r);
        double acc = 0.0;
        subwarp_spmv_any<true>(m.bt_lanes, m.bt_off, m.bt_idx, m.bt_val, m.bt_rows, t, [&](u32 row, double v) {
            const double rv = __dadd_rn(v, __dmul_rn(nc, qprev[row])); // svd_daxpy(-rnm, q_{j-1}, r), :1277
            rs[row] = rv;
            acc = fma(rv, x[row], acc);                                // alf = r . q_j, :1278
        });
        __syncthreads();
        const double tsum = cta_reduce(acc, sm);
        if (threadIdx.x == 0) work[blockIdx.x] = tsum;
        parts = work;
        nparts = (int)gridDim.x;
        coef_seen[0] = c.coef0;
        chain_barrier(bar, bar + 1);
        k0 = 1;
    }
    // Short vectors (SPMV launches only): ONE CTA runs the remaining phases, so that the 3-5 dependent reductions cost a block
    // barrier each instead of a grid barrier (a few us each with 148 CTAs); the other CTAs are done.
    const bool solo = SPMV && m.solo_tail;
    if (solo && blockIdx.x != 0) return;
    const u64 first = solo ? threadIdx.x : (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = solo ? blockDim.x : (u64)gridDim.x * blockDim.x;
    for (int k = k0; k < c.nphases; k++) {
        const double coef = (k == 0 && parts == nullptr) ? c.coef0 : cta_sum_partials(parts, nparts, &cs);
        coef_seen[k] = coef;
        const double nc = -coef;
        const double2 *__restrict__ x = reinterpret_cast<const double2 *>(c.x[k]);
        const double2 *__restrict__ y = reinterpret_cast<const double2 *>(c.y[k]);
        double acc = 0.0;
        for (u64 i = first; i < n2; i += stride) {
            // (right after the SpMV phase r holds what OTHER CTAs stored in this launch: read it through L2)
            double2 rv = (SPMV && k == k0) ? __ldcg(r + i) : r[i];
            const double2 xv = x[i];
            rv.x = __dadd_rn(rv.x, __dmul_rn(nc, xv.x)); // svd_daxpy(-c, x, r)
            rv.y = __dadd_rn(rv.y, __dmul_rn(nc, xv.y));
            r[i] = rv;
            const double2 yv = y ? (c.y[k] == c.x[k] ? xv : y[i]) : rv;
            acc = fma(rv.x, yv.x, acc);
            acc = fma(rv.y, yv.y, acc);
        }
        __syncthreads(); // sm / cs are reused by the next phase
        const double t = cta_reduce(acc, sm);
        double *out = work + (size_t)k * kMaxParts;
        if (threadIdx.x == 0) out[solo ? 0 : blockIdx.x] = t;
        parts = out;
        nparts = solo ? 1 : (int)gridDim.x;
        if (solo) __syncthreads();
        else chain_barrier(bar, bar + 1);
    }
    const double nrm2 = cta_sum_partials(parts, nparts, &cs);
    if (c.q_next) {
        // q_{j+1} = r / rnm exactly as the host-launched k_scale would form it (rnm = sqrt(|r|^2), alpha = 1 / rnm)
        const double alpha = 1.0 / sqrt(nrm2);
        double2 *__restrict__ qn = reinterpret_cast<double2 *>(c.q_next);
        for (u64 i = first; i < n2; i += stride) {
            double2 v = r[i];
            v.x = __dmul_rn(alpha, v.x);
            v.y = __dmul_rn(alpha, v.y);
            qn[i] = v;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        host_out[0] = coef_seen[c.pick[0]];
        host_out[1] = coef_seen[c.pick[1]];
        host_out[2] = coef_seen[c.pick[2]];
        host_out[3] = nrm2;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(host_out + 4) = seq;
    }
}

__global__ void __launch_bounds__(256) k_max_row_len(const u32 *__restrict__ off, u32 rows, u32 *__restrict__ out) {
    u32 best = 0;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) best = max(best, off[r + 1] - off[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out, best);
}

// out[i * n + c] = V[i * ld + old2new[c]]: the N-side singular vectors back in the caller's numbering (coalesced writes)
__global__ void __launch_bounds__(kVecThreads)
k_unpermute_rows(const double *__restrict__ V, u64 ld, u64 n, const u32 *__restrict__ old2new, double *__restrict__ out) {
    const double *v = V + (u64)blockIdx.y * ld;
    double *o = out + (u64)blockIdx.y * n;
    u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (; c < n; c += stride) o[c] = v[old2new[c]];
}

} // namespace

void unpermute_rows(const double *V, u64 ld, u64 n, u32 d, const u32 *old2new, double *out, cudaStream_t s) {
    if (d == 0 || n == 0) return;
    const unsigned gx = (unsigned)std::max<u64>(1, std::min<u64>((n + kVecThreads * 4 - 1) / (kVecThreads * 4), 4 * kNumSMs));
    for (u32 d0 = 0; d0 < d; d0 += 65535) {
        const u32 nd = std::min<u32>(65535, d - d0);
        k_unpermute_rows<<<dim3(gx, nd), kVecThreads, 0, s>>>(V + (u64)d0 * ld, ld, n, old2new, out + (u64)d0 * n);
    }
    LAS2_LAUNCH_CHECK();
}

unsigned step_chain_max_grid() {
    // one CTA per SM at most: every CTA of the launch is then resident at once, which the in-kernel barrier relies on
    static unsigned cached[kMaxDevicesTracked];
    static bool known[kMaxDevicesTracked];
    const int slot = current_device_slot();
    if (!known[slot]) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cached[slot] = (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) ? (unsigned)sms : 0u;
        cudaGetLastError();
        known[slot] = true;
    }
    return cached[slot];
}

void vec_step_chain(double *r, u64 ld, const StepChain &c, double *work, double *host_out, unsigned long long seq, cudaStream_t s) {
    u64 n2 = ld / 2;
    unsigned grid = std::min<unsigned>(step_chain_max_grid(), vec_grid(ld));
    if (grid == 0) throw Error(SVDLAS2_ERR_CUDA, "step chain: no device");
    // one CTA per SM at most (the in-kernel barrier): long vectors get 1024 threads per CTA to keep enough loads in flight
    const unsigned threads = n2 > (u64)grid * kVecThreads * 2 ? 1024u : (unsigned)kVecThreads;
    k_step_chain<false><<<grid, threads, 0, s>>>(reinterpret_cast<double2 *>(r), n2, c, StepSpmv{}, work, host_out, seq);
    LAS2_LAUNCH_CHECK();
}

int step_spmv_lanes(double mean_row_len) {
    int lanes = 2;
    while (lanes < 32 && (double)lanes < mean_row_len * 0.5) lanes <<= 1;
    return lanes;
}

u32 csr_max_row_len(const u32 *off, u32 rows, cudaStream_t s) {
    if (rows == 0) return 0;
    DevBuf<u32> out(1);
    LAS2_CUDA(cudaMemsetAsync(out.get(), 0, sizeof(u32), s));
    const unsigned grid = (unsigned)std::min<u64>(((u64)rows + 255) / 256, (u64)kNumSMs * 8);
    k_max_row_len<<<grid, 256, 0, s>>>(off, rows, out.get());
    LAS2_LAUNCH_CHECK();
    u32 h = 0;
    LAS2_CUDA(cudaMemcpyAsync(&h, out.get(), sizeof(u32), cudaMemcpyDeviceToHost, s));
    LAS2_CUDA(cudaStreamSynchronize(s));
    return h;
}

void vec_step_fused(double *r, u64 ld, const StepChain &c, const StepSpmv &m, double *work, double *host_out, unsigned long long seq,
                    cudaStream_t s) {
    const u64 n2 = ld / 2;
    const unsigned max_grid = step_chain_max_grid();
    if (max_grid == 0) throw Error(SVDLAS2_ERR_CUDA, "fused step: no device");
    // enough threads that every row of the larger product gets its sub-warp in about one pass (one CTA per SM at most)
    const u64 want = std::max<u64>({(u64)m.b_rows * (u64)m.b_lanes, (u64)m.bt_rows * (u64)m.bt_lanes, n2});
    const unsigned threads = want > (u64)max_grid * 256 ? 1024u : 256u;
    const unsigned grid = (unsigned)std::max<u64>(1, std::min<u64>(max_grid, (want + threads - 1) / threads));
    StepSpmv mm = m;
    mm.solo_tail = n2 <= 16384 ? 1 : 0; // <= 16 double2 per thread of one 1024-thread CTA
    k_step_chain<true><<<grid, mm.solo_tail ? 1024u : threads, 0, s>>>(reinterpret_cast<double2 *>(r), n2, c, mm, work, host_out, seq);
    LAS2_LAUNCH_CHECK();
}

unsigned vec_grid(u64 ld) {
    u64 n2 = ld / 2;
    u64 g = (n2 + (u64)kVecThreads * 2 - 1) / ((u64)kVecThreads * 2); // ~2 double2 per thread before striding
    if (g < 1) g = 1;
    if (g > (u64)kMaxParts) g = kMaxParts;
    return (unsigned)g;
}

void vec_scale(double *dst, const double *src, double alpha, u64 ld, cudaStream_t s) {
    k_scale<<<vec_grid(ld), kVecThreads, 0, s>>>((double2 *)dst, (const double2 *)src, alpha, ld / 2);
    LAS2_LAUNCH_CHECK();
}

void vec_dot(const double *a, const double *b, u64 ld, PartialSlot &out, cudaStream_t s) {
    unsigned g = vec_grid(ld);
    k_dot<<<g, kVecThreads, 0, s>>>((const double2 *)a, (const double2 *)b, ld / 2, out.parts);
    out.nparts = (int)g;
    LAS2_LAUNCH_CHECK();
}

void vec_axpy_dot(double *r, const double *x, Coef coef, const double *y, u64 ld, PartialSlot &out, cudaStream_t s) {
    unsigned g = vec_grid(ld);
    k_axpy_dot<true><<<g, kVecThreads, 0, s>>>((double2 *)r, (const double2 *)x, coef, (const double2 *)y, ld / 2,
                                               out.parts);
    out.nparts = (int)g;
    LAS2_LAUNCH_CHECK();
}

void vec_axpy(double *r, const double *x, Coef coef, u64 ld, cudaStream_t s) {
    k_axpy_dot<false><<<vec_grid(ld), kVecThreads, 0, s>>>((double2 *)r, (const double2 *)x, coef, nullptr, ld / 2,
                                                           nullptr);
    LAS2_LAUNCH_CHECK();
}

void vec_sumsq_any(const double *a, u64 n, PartialSlot &out, cudaStream_t s) {
    u64 g = (n + kVecThreads * 4 - 1) / (kVecThreads * 4);
    if (g < 1) g = 1;
    if (g > (u64)kMaxParts) g = kMaxParts;
    k_sumsq_any<<<(unsigned)g, kVecThreads, 0, s>>>(a, n, out.parts);
    out.nparts = (int)g;
    LAS2_LAUNCH_CHECK();
}

void vec_sigma_scale(double *u, u64 m, const PartialSlot &t, double *sigma_out, cudaStream_t s) {
    u64 g = (m + kVecThreads * 4 - 1) / (kVecThreads * 4);
    if (g < 1) g = 1;
    if (g > (u64)kMaxParts) g = kMaxParts;
    k_sigma_scale<<<(unsigned)g, kVecThreads, 0, s>>>(u, m, t.parts, t.nparts, sigma_out);
    LAS2_LAUNCH_CHECK();
}

void gather_scalars(const GatherList &g, double *out, cudaStream_t s) {
    k_gather_scalars<<<1, 32 * 8, 0, s>>>(g, out);
    LAS2_LAUNCH_CHECK();
}

} // namespace las2
