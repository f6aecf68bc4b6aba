// panel.cu -- blocked operations on the device-resident Lanczos-vector panel Q (N x J, column-major).
//
//  * panel_cgs2: the reorthogonalisation sweep of purge (reference src/lib.rs:1372-1388) and of the restart
//    path of startv (:1132-1135) as a tall-skinny GEMV pair  c = Q'[a b]  then  [a b] -= Q c.
//    The reference computes every coefficient against FROZEN copies (w3, w4) while it updates (w1, w0), i.e.
//    classical Gram-Schmidt, so "all dots first, then all updates" is the same arithmetic; the update keeps
//    the reference's order (k ascending, one rounded multiply and one rounded add per term).
//    Minimum traffic: 2 * ncols * N * 8 B (Q is read twice) + 6 * N * 8 B -- HBM-bound.
//  * ritz_contract: the Ritz back-transform V = Q Z_sel of ritvec (:1807-1817) on the FP64 FMA pipe.
#include "vecops.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "comm.hpp"

namespace las2 {

namespace {

constexpr int kPanelRows = 2048; // rows of (a, b) staged in shared memory per CTA
constexpr int kPanelThreads = 256;
constexpr int kPanelWarps = kPanelThreads / 32;

// partials[(rb * ncols + c) * NRHS + rhs] = sum over the row block rb of Q[i, col0 + c] * rhs[i]
template <int NRHS>
__global__ void __launch_bounds__(kPanelThreads)
k_panel_dots(const double *__restrict__ Q, u64 ld, u32 col0, u32 ncols, const double *__restrict__ a,
             const double *__restrict__ b, double *__restrict__ partials, u32 rb0) {
    __shared__ __align__(16) double sa[kPanelRows];
    __shared__ __align__(16) double sb[NRHS > 1 ? kPanelRows : 2];
    const u64 row0 = (u64)(blockIdx.x + rb0) * kPanelRows; // rb0: first row block of this rank's share
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kPanelRows; i += kPanelThreads) {
        const u64 r = row0 + i;
        sa[i] = (r < ld) ? a[r] : 0.0;
        if (NRHS > 1) sb[i] = (r < ld) ? b[r] : 0.0;
    }
    __syncthreads();
    const u32 rows_here = (u32)min((u64)kPanelRows, ld - row0); // multiple of 32 (ld is)
    const u32 wslots = gridDim.y * kPanelWarps;
    for (u32 c = blockIdx.y * kPanelWarps + warp; c < ncols; c += wslots) {
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(Q + (u64)(col0 + c) * ld + row0);
        double acc_a = 0.0, acc_b = 0.0;
#pragma unroll 8
        for (u32 i = lane; i < rows_here / 2; i += 32) {
            const double2 qv = q[i];
            const double2 av = *reinterpret_cast<const double2 *>(&sa[2 * i]);
            acc_a = fma(qv.x, av.x, acc_a);
            acc_a = fma(qv.y, av.y, acc_a);
            if (NRHS > 1) {
                const double2 bv = *reinterpret_cast<const double2 *>(&sb[2 * i]);
                acc_b = fma(qv.x, bv.x, acc_b);
                acc_b = fma(qv.y, bv.y, acc_b);
            }
        }
        acc_a = warp_sum(acc_a);
        if (NRHS > 1) acc_b = warp_sum(acc_b);
        if (lane == 0) {
            double *o = partials + ((u64)blockIdx.x * ncols + c) * NRHS;
            o[0] = acc_a;
            if (NRHS > 1) o[1] = acc_b;
        }
    }
}

// coeffs[rhs * ncols + c] = sum over row blocks ; abs_parts[blockIdx.x * 2 + rhs] = sum |coeff| of the CTA's columns.
// A CTA of 128 x 8 threads owns 128 columns; the row blocks are cut into 8 contiguous ranges (one per threadIdx.y), each summed
// in order, and the 8 range sums are added in range order: a fixed order, and an eighth of the dependent-add chain (977 row
// blocks at cfg 4: 185 us per sweep with one thread per column).
constexpr int kReduceSplit = 8;
template <int NRHS>
__global__ void __launch_bounds__(128 * kReduceSplit)
k_panel_reduce(const double *__restrict__ partials, u32 nblocks, u32 ncols, double *__restrict__ coeffs,
               double *__restrict__ abs_a, double *__restrict__ abs_b) {
    __shared__ double sm[2][4];
    __shared__ double part[kReduceSplit][NRHS][128];
    const u32 c = blockIdx.x * 128 + threadIdx.x;
    double ca = 0.0, cb = 0.0;
    {
        const u32 lo = (u32)((u64)nblocks * threadIdx.y / kReduceSplit), hi = (u32)((u64)nblocks * (threadIdx.y + 1) / kReduceSplit);
        double pa = 0.0, pb = 0.0;
        if (c < ncols) {
            for (u32 rb = lo; rb < hi; rb++) {
                const double *p = partials + ((u64)rb * ncols + c) * NRHS;
                pa += p[0];
                if (NRHS > 1) pb += p[1];
            }
        }
        part[threadIdx.y][0][threadIdx.x] = pa;
        if (NRHS > 1) part[threadIdx.y][NRHS - 1][threadIdx.x] = pb;
    }
    __syncthreads();
    if (threadIdx.y != 0) return;
    if (c < ncols) {
#pragma unroll
        for (int k = 0; k < kReduceSplit; k++) {
            ca += part[k][0][threadIdx.x];
            if (NRHS > 1) cb += part[k][NRHS - 1][threadIdx.x];
        }
        coeffs[c] = ca;
        if (NRHS > 1) coeffs[ncols + c] = cb;
    }
    double va = warp_sum(fabs(ca)), vb = warp_sum(fabs(cb));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = va; sm[1][warp] = vb; }
    // (only the 128 threads of threadIdx.y == 0 are left: a named barrier over exactly them)
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
        abs_a[blockIdx.x] = ((sm[0][0] + sm[0][1]) + sm[0][2]) + sm[0][3];
        abs_b[blockIdx.x] = ((sm[1][0] + sm[1][1]) + sm[1][2]) + sm[1][3];
    }
}

constexpr int kUpdThreads = 128;
constexpr int kUpdChunk = 512; // coefficients staged per pass

// a -= Q ca ; b -= Q cb (k ascending, mul then add, as the reference's chain of daxpys) ; partial of a . b
template <int NRHS>
__global__ void __launch_bounds__(kUpdThreads)
k_panel_update(const double *__restrict__ Q, u64 ld, u32 col0, u32 ncols, const double *__restrict__ coeffs,
               double *__restrict__ a, double *__restrict__ b, double *__restrict__ dot_parts, u64 i0, u64 i1) {
    __shared__ double sc[NRHS][kUpdChunk];
    __shared__ double sm[kUpdThreads / 32];
    const u64 i = i0 + (u64)blockIdx.x * kUpdThreads + threadIdx.x; // double2 index inside this rank's share [i0, i1)
    const bool live = i < i1;
    double2 av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
    if (live) {
        av = reinterpret_cast<double2 *>(a)[i];
        if (NRHS > 1) bv = reinterpret_cast<double2 *>(b)[i];
    }
    for (u32 c0 = 0; c0 < ncols; c0 += kUpdChunk) {
        const u32 nc = min((u32)kUpdChunk, ncols - c0);
        __syncthreads();
        for (u32 k = threadIdx.x; k < nc; k += kUpdThreads) {
            sc[0][k] = -coeffs[c0 + k];
            if (NRHS > 1) sc[1][k] = -coeffs[ncols + c0 + k];
        }
        __syncthreads();
        if (live) {
            const double2 *__restrict__ q = reinterpret_cast<const double2 *>(Q + (u64)(col0 + c0) * ld) + i;
            const u64 ld2 = ld / 2;
#pragma unroll 8
            for (u32 k = 0; k < nc; k++) {
                const double2 qv = q[(u64)k * ld2];
                av.x = __dadd_rn(av.x, __dmul_rn(sc[0][k], qv.x));
                av.y = __dadd_rn(av.y, __dmul_rn(sc[0][k], qv.y));
                if (NRHS > 1) {
                    bv.x = __dadd_rn(bv.x, __dmul_rn(sc[1][k], qv.x));
                    bv.y = __dadd_rn(bv.y, __dmul_rn(sc[1][k], qv.y));
                }
            }
        }
    }
    double acc = 0.0;
    if (live) {
        reinterpret_cast<double2 *>(a)[i] = av;
        if (NRHS > 1) {
            reinterpret_cast<double2 *>(b)[i] = bv;
            acc = fma(av.x, bv.x, 0.0);
            acc = fma(av.y, bv.y, acc);
        }
    }
    if (NRHS > 1 && dot_parts) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        acc = warp_sum(acc);
        if (lane == 0) sm[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) dot_parts[blockIdx.x] = ((sm[0] + sm[1]) + sm[2]) + sm[3];
    }
}

// The same update for SHORT vectors.  With one double2 per thread a vector of 10,000 rows is 39 CTAs of the kernel above, and
// the ~8 loads a thread keeps in flight are far too few bytes for the memory latency: 85 us per sweep of ~400 columns at cfg 1
// (profiles/r2v_launches_cfg1_summary.csv; the compiler does not hoist more loads whatever the unroll factor).  Here every
// thread streams its element of the next 24-48 columns into shared memory with cp.async (4 stages of 16 or 8 columns, no
// registers held, no barrier: a thread only ever reads what it copied itself), CTAs of 32 or 64 threads spread the rows over
// two to four times as many SMs, and the arithmetic -- order included -- is unchanged.
constexpr int kDeepStages = 4; // x (16 columns x 32 threads) or (8 columns x 64 threads) x 16 B = 32 KB of static shared memory
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NRHS, int kDeepThreads, int kDeepCols>
__global__ void __launch_bounds__(kDeepThreads)
k_panel_update_deep(const double *__restrict__ Q, u64 ld, u32 col0, u32 ncols, const double *__restrict__ coeffs,
                    double *__restrict__ a, double *__restrict__ b, double *__restrict__ dot_parts, u64 i0, u64 i1) {
    __shared__ double sc[NRHS][kUpdChunk];
    __shared__ double sm[2];
    __shared__ __align__(16) double2 stage[kDeepStages][kDeepCols][kDeepThreads];
    const u64 i = i0 + (u64)blockIdx.x * kDeepThreads + threadIdx.x;
    const bool live = i < i1;
    double2 av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
    if (live) {
        av = reinterpret_cast<double2 *>(a)[i];
        if (NRHS > 1) bv = reinterpret_cast<double2 *>(b)[i];
    }
    const u64 ld2 = ld / 2;
    for (u32 c0 = 0; c0 < ncols; c0 += kUpdChunk) {
        const u32 nc = min((u32)kUpdChunk, ncols - c0);
        __syncthreads();
        for (u32 k = threadIdx.x; k < nc; k += kDeepThreads) {
            sc[0][k] = -coeffs[c0 + k];
            if (NRHS > 1) sc[1][k] = -coeffs[ncols + c0 + k];
        }
        __syncthreads();
        if (!live) continue;
        const double2 *__restrict__ q = reinterpret_cast<const double2 *>(Q + (u64)(col0 + c0) * ld) + i;
        const u32 ngroups = (nc + kDeepCols - 1) / kDeepCols;
        auto issue = [&](u32 g) { // columns [g * 16, g * 16 + 16) of this chunk into slot g % stages (always one commit)
            if (g < ngroups) {
                const u32 k0 = g * kDeepCols, kn = min((u32)kDeepCols, nc - k0);
                for (u32 u = 0; u < kn; u++) cp_async16(&stage[g % kDeepStages][u][threadIdx.x], q + (u64)(k0 + u) * ld2);
            }
            cp_async_commit();
        };
        for (u32 g = 0; g + 1 < (u32)kDeepStages; g++) issue(g);
        for (u32 g = 0; g < ngroups; g++) {
            issue(g + kDeepStages - 1);
            cp_async_wait<kDeepStages - 1>(); // group g has landed
            const u32 k0 = g * kDeepCols, kn = min((u32)kDeepCols, nc - k0);
#pragma unroll 4
            for (u32 u = 0; u < kn; u++) {
                const double2 qv = stage[g % kDeepStages][u][threadIdx.x];
                const u32 k = k0 + u;
                av.x = __dadd_rn(av.x, __dmul_rn(sc[0][k], qv.x));
                av.y = __dadd_rn(av.y, __dmul_rn(sc[0][k], qv.y));
                if (NRHS > 1) {
                    bv.x = __dadd_rn(bv.x, __dmul_rn(sc[1][k], qv.x));
                    bv.y = __dadd_rn(bv.y, __dmul_rn(sc[1][k], qv.y));
                }
            }
        }
        cp_async_wait<0>();
    }
    double acc = 0.0;
    if (live) {
        reinterpret_cast<double2 *>(a)[i] = av;
        if (NRHS > 1) {
            reinterpret_cast<double2 *>(b)[i] = bv;
            acc = fma(av.x, bv.x, 0.0);
            acc = fma(av.y, bv.y, acc);
        }
    }
    if (NRHS > 1 && dot_parts) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        acc = warp_sum(acc);
        if (lane == 0) sm[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) dot_parts[blockIdx.x] = kDeepThreads > 32 ? sm[0] + sm[1] : sm[0];
    }
}

// abs_a[cta] / abs_b[cta] = sum |coeff| over the CTA's 128 columns (sharded sweep: the coefficients are complete only
// after the all-reduce, so the sums cannot come out of k_panel_reduce)
template <int NRHS>
__global__ void __launch_bounds__(128)
k_coeff_abs(const double *__restrict__ coeffs, u32 ncols, double *__restrict__ abs_a, double *__restrict__ abs_b) {
    __shared__ double sm[2][4];
    const u32 c = blockIdx.x * 128 + threadIdx.x;
    const double ca = c < ncols ? coeffs[c] : 0.0;
    const double cb = (NRHS > 1 && c < ncols) ? coeffs[ncols + c] : 0.0;
    double va = warp_sum(fabs(ca)), vb = warp_sum(fabs(cb));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sm[0][warp] = va; sm[1][warp] = vb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        abs_a[blockIdx.x] = ((sm[0][0] + sm[0][1]) + sm[0][2]) + sm[0][3];
        abs_b[blockIdx.x] = ((sm[1][0] + sm[1][1]) + sm[1][2]) + sm[1][3];
    }
}

// second-stage for dot partials when the update grid exceeds kMaxParts: fold groups of partials in order
__global__ void k_fold_partials(const double *__restrict__ in, u32 n, u32 group, double *__restrict__ out) {
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    u32 lo = g * group;
    if (lo >= n) return;
    u32 hi = min(lo + group, n);
    double s = 0.0;
    for (u32 i = lo; i < hi; i++) s += in[i];
    out[g] = s;
}

// ---- Ritz contraction ---------------------------------------------------------------------------------
constexpr int kRitzThreads = 128;
constexpr int kRitzCols = 8;   // output columns per thread (x2 rows = 16 accumulators)
constexpr int kRitzChunk = 256;

__global__ void __launch_bounds__(kRitzThreads)
k_ritz_contract(const double *__restrict__ Q, u64 ld, u32 js, const double *__restrict__ C, u32 d,
                double *__restrict__ V, u64 i0, u64 i1) {
    __shared__ double sc[kRitzChunk][kRitzCols];
    const u64 i = i0 + (u64)blockIdx.x * kRitzThreads + threadIdx.x; // double2 row index inside this rank's share [i0, i1)
    const u32 c0 = blockIdx.y * kRitzCols;
    const bool live = i < i1;
    double2 acc[kRitzCols];
#pragma unroll
    for (int t = 0; t < kRitzCols; t++) acc[t] = make_double2(0.0, 0.0);
    const u64 ld2 = ld / 2;
    for (u32 k0 = 0; k0 < js; k0 += kRitzChunk) {
        const u32 nk = min((u32)kRitzChunk, js - k0);
        __syncthreads();
        for (u32 e = threadIdx.x; e < nk * kRitzCols; e += kRitzThreads) {
            const u32 k = e / kRitzCols, t = e % kRitzCols;
            sc[k][t] = (c0 + t < d) ? C[(u64)(k0 + k) * d + c0 + t] : 0.0;
        }
        __syncthreads();
        if (live) {
            const double2 *__restrict__ q = reinterpret_cast<const double2 *>(Q + (u64)k0 * ld) + i;
#pragma unroll 4
            for (u32 k = 0; k < nk; k++) {
                const double2 qv = q[(u64)k * ld2];
#pragma unroll
                for (int t = 0; t < kRitzCols; t++) {
                    acc[t].x = fma(sc[k][t], qv.x, acc[t].x);
                    acc[t].y = fma(sc[k][t], qv.y, acc[t].y);
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int t = 0; t < kRitzCols; t++)
            if (c0 + t < d) reinterpret_cast<double2 *>(V + (u64)(c0 + t) * ld)[i] = acc[t];
    }
}

// ---- Ritz contraction on the FP64 tensor path (DMMA m8n8k4) ----------------------------------------------
// V[rows, c0 .. c0+64) = Q[rows, 0 .. js) . C[0 .. js, c0 .. c0+64) for a block of 128 rows: 8 warps, each a 32 x 32 tile of
// 4 x 4 mma tiles (32 accumulators per thread); the 128 x 16 slice of Q and the 16 x 64 slice of C of every k-step are staged
// in shared memory (double-buffered, cp.async), so a row block of Q is read d / 64 times instead of d / 8 (k_ritz_contract re-reads
// the panel 125 times at d = 1000 and is bandwidth-bound on exactly that, profiles/r2g_ncu_summary.md section 2).  Row padding of 4
// doubles makes both fragment reads conflict-free.  The accumulation order is fixed (k ascending, 4 products per instruction).
constexpr int kDmTM = 128, kDmTN = 64, kDmTK = 16, kDmThreads = 256;
constexpr int kDmLdA = kDmTM + 4, kDmLdB = kDmTN + 4;

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = pred ? 16 : 0; // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kDmThreads, 2)
k_ritz_contract_dmma(const double *__restrict__ Q, u64 ld, u32 js, const double *__restrict__ C, u32 d, double *__restrict__ V,
                     u64 row_lo, u64 row_hi) {
    extern __shared__ __align__(16) double dm_smem[];
    double (*As)[kDmTK][kDmLdA] = reinterpret_cast<double (*)[kDmTK][kDmLdA]>(dm_smem);
    double (*Bs)[kDmTK][kDmLdB] = reinterpret_cast<double (*)[kDmTK][kDmLdB]>(dm_smem + 2 * kDmTK * kDmLdA);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;       // warp tile: rows wm*32, cols wn*32
    const int fr = lane >> 2, fk = lane & 3;        // fragment coordinates
    const u64 row0 = row_lo + (u64)blockIdx.x * kDmTM;
    const u32 c0 = blockIdx.y * kDmTN;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_tiles = [&](int buf, u32 k0) {
        // Q slice: 16 k x 128 rows = 1024 x 16 B, four per thread (rows are contiguous in the column-major panel)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int e = tid + t * kDmThreads;       // 0 .. 1023
            const int k = e >> 6, r2 = (e & 63) * 2;  // row pair
            const u64 row = row0 + r2;
            const bool ok = (k0 + k < js) && (row < row_hi);
            const double *src = ok ? Q + (u64)(k0 + k) * ld + row : Q;
            cp_async16(&As[buf][k][r2], src, ok);
        }
        // C slice: 16 k x 64 columns (row-major js x d, any d: plain loads with guards)
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int e = tid + t * kDmThreads;       // 0 .. 1023
            const int k = e >> 6, c = e & 63;
            const bool ok = (k0 + k < js) && (c0 + c < d);
            Bs[buf][k][c] = ok ? C[(u64)(k0 + k) * d + c0 + c] : 0.0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const u32 nk = (js + kDmTK - 1) / kDmTK;
    load_tiles(0, 0);
    for (u32 it = 0; it < nk; it++) {
        const int buf = it & 1;
        if (it + 1 < nk) {
            load_tiles(buf ^ 1, (it + 1) * kDmTK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kDmTK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[buf][kk + fk][wm * 32 + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = Bs[buf][kk + fk][wn * 32 + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
    // C fragment: row = lane / 4, columns (lane % 4) * 2 and + 1
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u64 row = row0 + wm * 32 + i * 8 + fr;
        if (row >= row_hi) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const u32 col = c0 + wn * 32 + j * 8 + fk * 2;
            if (col < d) V[(u64)col * ld + row] = acc[i][j][0];
            if (col + 1 < d) V[(u64)(col + 1) * ld + row] = acc[i][j][1];
        }
    }
}

} // namespace

void panel_cgs2(const double *Q, u64 ld, u64 n, u32 col0, u32 ncols, double *a, double *b, PanelWork &work,
                PartialSlot *abs_a, PartialSlot *abs_b, PartialSlot *dot_out, cudaStream_t s, u64 *launches,
                cudaEvent_t ev_begin, cudaEvent_t ev_mid, cudaEvent_t ev_end, Comm *comm, u64 *sharded_sweeps) {
    (void)n;
    if (ncols == 0) return;
    const u32 nblocks_all = cdiv(ld, kPanelRows);
    // split the rows over the ranks when the sweep is big enough (same decision on every rank: sizes and environment only)
    static const double shard_mb = [] { const char *e = std::getenv("SVDLAS2_SHARD_PURGE_MB"); return e ? std::atof(e) : 512.0; }();
    const int G = comm ? comm->world() : 1;
    const bool shard = G > 1 && nblocks_all >= (u32)G && 2.0 * ncols * (double)ld * 8.0 >= shard_mb * 1048576.0;
    u32 rb0 = 0, rb1 = nblocks_all;
    std::vector<u64> offs;
    if (shard) {
        const int g = comm->rank();
        rb0 = (u32)((u64)nblocks_all * g / G);
        rb1 = (u32)((u64)nblocks_all * (g + 1) / G);
        offs.resize(G + 1);
        for (int k = 0; k <= G; k++) offs[k] = std::min<u64>((u64)nblocks_all * k / G * kPanelRows, ld);
        if (sharded_sweeps) *sharded_sweeps += 1;
    }
    const u32 nblocks = rb1 - rb0;
    const u64 row_lo = std::min<u64>((u64)rb0 * kPanelRows, ld), row_hi = std::min<u64>((u64)rb1 * kPanelRows, ld);
    const u64 i0 = row_lo / 2, i1 = row_hi / 2; // double2 units (ld and the block size are even)
    if (work.cap_cols < ncols || work.cap_blocks < nblocks) {
        size_t cc = work.cap_cols < ncols ? (size_t)ncols * 2 : work.cap_cols;
        size_t cb = std::max<size_t>(work.cap_blocks, nblocks);
        work.partials.alloc(cb * cc * 2);
        work.coeffs.alloc(cc * 2);
        work.cap_cols = cc;
        work.cap_blocks = cb;
    }
    u32 cs = cdiv(kMaxParts, std::max(1u, nblocks));
    const u32 max_cs = cdiv(ncols, kPanelWarps);
    if (cs > max_cs) cs = max_cs;
    if (cs < 1) cs = 1;
    const u32 nred = cdiv(ncols, 128);
    if (nred > (u32)kMaxParts) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "panel_cgs2: too many columns");
    dim3 grid(nblocks, cs);
    // short vectors (fewer than two CTAs of the plain kernel per SM): the cp.async-pipelined sibling on CTAs of 64 threads
    static const bool deep_on = [] { const char *e = std::getenv("SVDLAS2_PANEL_DEEP"); return !(e && e[0] == '0'); }();
    const bool deep = deep_on && cdiv(i1 - i0, kUpdThreads) < 2u * (u32)kNumSMs;
    const u32 deep_threads = (i1 - i0) <= 32u * (u32)kMaxParts ? 32u : 64u; // (one dot partial per CTA, at most kMaxParts of them)
    const u32 ugrid = cdiv(i1 - i0, deep ? deep_threads : (u32)kUpdThreads);
    const int nrhs = b ? 2 : 1;
    if (ev_begin) LAS2_CUDA(cudaEventRecord(ev_begin, s)); // tuning hook (svdlas2_test_time_panel)
    // ---- coefficients ----
    if (b) {
        k_panel_dots<2><<<grid, kPanelThreads, 0, s>>>(Q, ld, col0, ncols, a, b, work.partials, rb0);
        k_panel_reduce<2><<<nred, dim3(128, kReduceSplit), 0, s>>>(work.partials, nblocks, ncols, work.coeffs, abs_a->parts, abs_b->parts);
    } else {
        k_panel_dots<1><<<grid, kPanelThreads, 0, s>>>(Q, ld, col0, ncols, a, nullptr, work.partials, rb0);
        k_panel_reduce<1><<<nred, dim3(128, kReduceSplit), 0, s>>>(work.partials, nblocks, ncols, work.coeffs, abs_a->parts, abs_b->parts);
    }
    if (launches) *launches += 2;
    if (shard) {
        comm->allreduce_sum(work.coeffs, (u64)nrhs * ncols, s);
        if (b) k_coeff_abs<2><<<nred, 128, 0, s>>>(work.coeffs, ncols, abs_a->parts, abs_b->parts);
        else k_coeff_abs<1><<<nred, 128, 0, s>>>(work.coeffs, ncols, abs_a->parts, abs_b->parts);
        if (launches) *launches += 1;
    }
    abs_a->nparts = abs_b->nparts = (int)nred;
    if (ev_mid) LAS2_CUDA(cudaEventRecord(ev_mid, s));
    // ---- update ----
    if (b) {
        // dot partials: one per update CTA, folded to <= kMaxParts (replicated sweep only; a sharded sweep takes the
        // dot product of the re-assembled pair instead)
        double *dp = (dot_out && !shard) ? dot_out->parts : nullptr;
        if (dot_out && !shard && ugrid > (u32)kMaxParts) {
            if (work.dotparts.size() < ugrid) work.dotparts.alloc(ugrid);
            dp = work.dotparts;
        }
        if (ugrid && deep && deep_threads == 32) k_panel_update_deep<2, 32, 16><<<ugrid, 32, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, b, dp, i0, i1);
        else if (ugrid && deep) k_panel_update_deep<2, 64, 8><<<ugrid, 64, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, b, dp, i0, i1);
        else if (ugrid) k_panel_update<2><<<ugrid, kUpdThreads, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, b, dp, i0, i1);
        if (launches) *launches += 1;
        if (dot_out && !shard) {
            if (ugrid > (u32)kMaxParts) {
                const u32 group = cdiv(ugrid, kMaxParts);
                const u32 ng = cdiv(ugrid, group);
                k_fold_partials<<<cdiv(ng, 128), 128, 0, s>>>(dp, ugrid, group, dot_out->parts);
                dot_out->nparts = (int)ng;
                if (launches) *launches += 1;
            } else {
                dot_out->nparts = (int)ugrid;
            }
        }
    } else {
        if (ugrid && deep && deep_threads == 32) k_panel_update_deep<1, 32, 16><<<ugrid, 32, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, nullptr, nullptr, i0, i1);
        else if (ugrid && deep) k_panel_update_deep<1, 64, 8><<<ugrid, 64, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, nullptr, nullptr, i0, i1);
        else if (ugrid) k_panel_update<1><<<ugrid, kUpdThreads, 0, s>>>(Q, ld, col0, ncols, work.coeffs, a, nullptr, nullptr, i0, i1);
        if (launches) *launches += 1;
    }
    if (shard) {
        comm->allgather_ranges2(a, b, offs.data(), s); // the updated (q, r) pair in ONE collective
        if (b && dot_out) {
            vec_dot(a, b, ld, *dot_out, s);
            if (launches) *launches += 1;
        }
    }
    if (ev_end) LAS2_CUDA(cudaEventRecord(ev_end, s));
    LAS2_LAUNCH_CHECK();
}

int g_ritz_variant = -1; // test / tuning hook (svdlas2_test_time_ritz): -1 auto, 0 FMA pipe, 1 DMMA

void ritz_contract(const double *Q, u64 ld, u64 n, u32 js, const double *C, u32 d, double *V, cudaStream_t s,
                   u64 *launches, Comm *comm, bool *sharded) {
    (void)n;
    if (sharded) *sharded = false;
    if (d == 0 || js == 0) return;
    // Sharded runs hold the panel replicated: for a contraction big enough to matter every rank forms only its row blocks
    // of V and one all-gather per Ritz vector completes them.  Each element is computed by exactly one rank with the
    // arithmetic of the replicated kernel, so the result is bit-identical to the single-GPU one on every rank.
    static const double shard_gflop = [] { const char *e = std::getenv("SVDLAS2_SHARD_RITZ_GFLOP"); return e ? std::atof(e) : 50.0; }();
    const int G = comm ? comm->world() : 1;
    const u32 nblocks_all = cdiv(ld, kPanelRows);
    const bool shard = G > 1 && nblocks_all >= (u32)G && 2.0 * (double)ld * js * d >= shard_gflop * 1e9;
    u64 i0 = 0, i1 = ld / 2;
    std::vector<u64> offs;
    if (shard) {
        offs.resize(G + 1);
        for (int k = 0; k <= G; k++) offs[k] = std::min<u64>((u64)nblocks_all * k / G * kPanelRows, ld);
        i0 = offs[comm->rank()] / 2;
        i1 = offs[comm->rank() + 1] / 2;
    }
    if (i1 > i0) {
        // SVDLAS2_RITZ: "dmma" (default for contractions whose panel re-reads would dominate) or "fma" (the 2 x 8 register
        // tile on the FMA pipe); both are bit-reproducible, each element is formed by one thread / one warp
        static const int want = [] { const char *e = std::getenv("SVDLAS2_RITZ"); return !e ? -1 : (e[0] == 'f' ? 0 : 1); }();
        const int variant = g_ritz_variant >= 0 ? g_ritz_variant : want;
        const bool dmma = variant == 1 || (variant < 0 && d >= 16 && js >= 16);
        if (dmma) {
            constexpr size_t smem = (size_t)2 * kDmTK * (kDmLdA + kDmLdB) * sizeof(double); // 51,200 B
            static bool configured[kMaxDevicesTracked];
            bool &done = configured[current_device_slot()];
            if (!done) {
                LAS2_CUDA(cudaFuncSetAttribute(k_ritz_contract_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                done = true;
            }
            dim3 grid(cdiv(2 * (i1 - i0), kDmTM), cdiv(d, kDmTN));
            k_ritz_contract_dmma<<<grid, kDmThreads, smem, s>>>(Q, ld, js, C, d, V, 2 * i0, 2 * i1);
        } else {
            dim3 grid(cdiv(i1 - i0, kRitzThreads), cdiv(d, kRitzCols));
            k_ritz_contract<<<grid, kRitzThreads, 0, s>>>(Q, ld, js, C, d, V, i0, i1);
        }
        LAS2_LAUNCH_CHECK();
        if (launches) *launches += 1;
    }
    if (shard) {
        for (u32 c = 0; c < d; c += 8) { // eight Ritz vectors per collective
            double *v[8];
            const int nv = (int)std::min<u32>(8, d - c);
            for (int k = 0; k < nv; k++) v[k] = V + (u64)(c + k) * ld;
            comm->allgather_ranges_n(v, nv, offs.data(), s);
        }
        if (sharded) *sharded = true;
    }
}

} // namespace las2
