// las2.cu -- the LAS2 driver: host control flow of lanso / lanczos_step / purge / ritvec, every O(N) or
// O(nnz) operation a kernel launch on the operator's stream.
//
// Reference call tree (src/lib.rs): svdLAS2 :570-655 -> lanso :1925-2017 -> stpone :1179-1201 / startv
// :1098-1147 -> lanczos_step :1233-1319 -> ortbnd :1431-1453, purge :1356-1400 -> imtqlb :962-1068,
// error_bound :1480-1532 -> ritvec :1767-1879 -> imtql2 :1576-1693.
//
// Device-resident state (replaces WorkSpace :740-779 and Store :707-737):
//   Q      N x J column-major panel of Lanczos vectors, leading dimension ld = roundup(N, 32).  Column j IS
//          the reference's w1 at step j (q_j); w2 is column j-1.  The reference's p-vectors (w3, w4 and
//          storp) are bitwise copies of the q-vectors at every point they are read (datx and dscal compute the
//          same products, :903-914), so they need no storage here.
//   r      the residual (w0; w4 is its copy taken right before every norm)
//   t      M-vector between the two SpMVs of opb (temp)
// Scalars (alf, bet, rnm, reorthogonalisation coefficients) are produced as per-CTA partials, consumed on the
// device by the next kernel of the chain, and read back ONCE per Lanczos step (one extra time per purge pass)
// for the host-side recurrences (ortbnd, the purge trigger, tol).
//
// How a step is launched: large operators -- the SpMV pair of opb (chunk_spmv.cu; sharded: + the exchange kernel of
// peer_ring.cu), then either the chain of k_axpy_dot kernels or ONE k_step_chain launch (vectors of up to 8 MB); small operators
// in plain CSR on one GPU -- the whole step incl. both SpMVs as one k_step_chain<true> launch (vecops.cuh, StepSpmv).
// Host recurrences: the analyses of T run imtqlb; the last one also records imtql2's rotation plan, which ritvec replays on
// the GPU (tridiag_device.cu) without running the O(js^2) chain again.
#include "las2.hpp"

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <future>
#include <mutex>
#include <new>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "host_pool.hpp"
#include "spmv.cuh"
#include "tridiag_device.cuh"

namespace las2 {

// vecops.cu
void vec_sigma_scale(double *u, u64 m, const PartialSlot &t, double *sigma_out, cudaStream_t s);
void vec_sumsq_any(const double *a, u64 n, PartialSlot &out, cudaStream_t s);
void unpermute_rows(const double *V, u64 ld, u64 n, u32 d, const u32 *old2new, double *out, cudaStream_t s);

namespace {

constexpr u64 kMaxLL = 2; // MAXLL, src/lib.rs:665

double wall() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

u64 round_up(u64 v, u64 m) { return (v + m - 1) / m * m; }

// The rotation plan an analysis records is O(js^2) doubles (86 MB at js 2252); growing a fresh vector to that size costs more
// than the recording itself (page faults + reallocation copies: +0.13 s at cfg 5, measured on the host alone), so one spare
// plan's storage is kept per process between solves, like the result and device buffers (svdlas2_trim_caches() drops it).
struct PlanSpare {
    std::mutex mu;
    QlPlan plan;
};
PlanSpare &plan_spare() {
    static PlanSpare *p = new PlanSpare(); // leaked on purpose, like the other caches
    return *p;
}

struct Solver {
    Operator &op;
    cudaStream_t s;
    const u64 N, M, ld;
    const u64 dims, iters;
    const double eps1; // eps * sqrt(N): roundoff estimate for the dot product of two unit vectors

    // device
    DevBuf<double> Q;
    u64 qcap = 0, qused = 0;
    DevBuf<double> r_own, xs, t;
    struct RPtr { double *p = nullptr; double *get() const { return p; } } r; // w0: the ring's buffer in a sharded run
    DevBuf<double> parts;
    PartialSlot slot[8];
    PinnedBuf<double> hs;
    PanelWork pw;
    // the step chain (vecops.cuh): partial scratch, call counter, and the panel column the chain has already filled with
    // the next Lanczos vector (none: ~0)
    DevBuf<double> chain_work;
    unsigned long long chain_seq = 0;
    u64 spec_col = ~(u64)0;
    bool use_chain = false;
    // small operators (plain CSR, short rows, one GPU): the two SpMVs of svd_opb run inside the chain's launch as well
    bool use_fused = false;
    StepSpmv fused{};
    // host (WorkSpace arrays, src/lib.rs:770-775)
    std::vector<double> alf, bet, eta, oldeta, ritz, bnd, w5;
    svdlas2_profile prof;
    u64 n_sharded_sweeps = 0; // purge / restart sweeps whose rows were split over the ranks
    double t_purge_host = 0.0, t_analysis_host = 0.0; // SVDLAS2_TRACE: host time inside purge (incl. its syncs) / the analyses of T
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> opb_events;
    // result buffers (pinned, from the host pool) are acquired on a helper thread while the GPU runs lanso
    cudaStream_t copy_stream = nullptr;
    std::future<void *> fut_m, fut_n; // dims x M and dims x N doubles
    bool prefetched = false;

    Solver(Operator &o, u64 dimensions, u64 iterations)
        : op(o), s(o.stream), N(o.N), M(o.M), ld(round_up(o.N, 32)), dims(dimensions), iters(iterations),
          eps1(kEps * std::sqrt((double)o.N)) {
        std::memset(&prof, 0, sizeof prof);
        {
            PlanSpare &ps = plan_spare();
            std::lock_guard<std::mutex> lk(ps.mu);
            kept_plan = std::move(ps.plan);
            ps.plan = QlPlan();
        }
        if (op.ring) { r.p = op.ring->r_mine(); } else { r_own.alloc(ld); r.p = r_own.get(); }
        xs.alloc(ld); t.alloc(M ? M : 1);
        LAS2_CUDA(cudaMemsetAsync(r.get(), 0, ld * sizeof(double), s));
        LAS2_CUDA(cudaMemsetAsync(xs.get(), 0, ld * sizeof(double), s));
        parts.alloc(8 * (size_t)kMaxParts);
        for (int i = 0; i < 8; i++) slot[i] = PartialSlot{parts.get() + (size_t)i * kMaxParts, 0};
        hs.alloc(16);
        for (int i = 0; i < 16; i++) hs[i] = 0.0;
        {
            const char *e = std::getenv("SVDLAS2_STEP_CHAIN");
            // default: for vectors of up to 8 MB, where the step is launch-latency-bound (cfg 1-3, cfg 5); the separate kernels
            // stream longer vectors faster than one CTA per SM can (SVDLAS2_STEP_CHAIN=0 / 1 forces it off / on)
            static const u64 max_ld = [] { const char *m = std::getenv("SVDLAS2_STEP_CHAIN_MAX_N"); return m ? (u64)std::atoll(m) : (u64)1 << 20; }();
            use_chain = step_chain_max_grid() > 0 && (e ? e[0] == '1' : ld <= max_ld);
            if (use_chain) {
                chain_work.alloc(6 * (size_t)kMaxParts + 1);
                LAS2_CUDA(cudaMemsetAsync(chain_work.get(), 0, (6 * (size_t)kMaxParts + 1) * sizeof(double), s));
            }
            const char *f = std::getenv("SVDLAS2_STEP_FUSED");
            if (use_chain && !(f && f[0] == '0') && !op.ring && !op.comm && !op.profile && !op.B.chk.active && !op.BT.chk.active &&
                M > 0 && N > 0 && op.B.nnz > 0 && M < (1ull << 31) && N < (1ull << 31)) {
                auto lanes_for = [&](const DeviceCsr &A, u64 rows, int *lanes) {
                    const u32 longest = csr_max_row_len(A.off.get(), (u32)rows, s);
                    int l = step_spmv_lanes((double)A.nnz / (double)rows);
                    while (l < 32 && longest > step_spmv_max_row(l)) l <<= 1;
                    *lanes = l;
                    return longest <= step_spmv_max_row(l);
                };
                int lb = 0, lbt = 0;
                if (lanes_for(op.B, M, &lb) && lanes_for(op.BT, N, &lbt)) {
                    use_fused = true;
                    fused = StepSpmv{op.B.off.get(), op.B.idx.get(), op.B.val.get(), (u32)M, lb,
                                     op.BT.off.get(), op.BT.idx.get(), op.BT.val.get(), (u32)N, lbt, t.get(), 0};
                }
            }
        }
        alf.assign(iters, 0.0); eta.assign(iters, 0.0); oldeta.assign(iters, 0.0);
        bet.assign(iters + 1, 0.0); ritz.assign(iters + 1, 0.0);
        bnd.assign(iters + 1, 1.7976931348623157e308); // f64::MAX, :775
        w5.assign(iters + 2, 0.0);
        LAS2_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        const u64 want = dims * (M + N) * sizeof(double);
        if (want <= host_prefetch_limit() && !std::getenv("SVDLAS2_NO_PREFETCH")) { // only when the upper bound d <= dims is affordable
            const int dev = op.device;
            u64 vb0 = 0, ve0 = 0;
            op.nside_rows(dims, &vb0, &ve0); // upper bound: d <= dims (a block of the dealt-out rows never exceeds this)
            const u64 vrows = op.nside_mode == 2 && op.comm ? (dims + op.comm->world() - 1) / op.comm->world() + 1 : ve0 - vb0;
            const size_t bm = std::max<u64>(1, dims * M) * sizeof(double), bn = std::max<u64>(1, vrows * N) * sizeof(double);
            fut_m = std::async(std::launch::async, [dev, bm] { cudaSetDevice(dev); return host_acquire(bm); });
            fut_n = std::async(std::launch::async, [dev, bn] { cudaSetDevice(dev); return host_acquire(bn); });
            prefetched = true;
        }
    }
    ~Solver() {
        {
            PlanSpare &ps = plan_spare();
            std::lock_guard<std::mutex> lk(ps.mu);
            if (kept_plan.sc.capacity() > ps.plan.sc.capacity()) ps.plan = std::move(kept_plan);
        }
        for (auto &e : opb_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        // buffers nobody took (an error before ritvec) go back to the pool
        try { if (fut_m.valid()) host_release(fut_m.get()); } catch (...) {}
        try { if (fut_n.valid()) host_release(fut_n.get()); } catch (...) {}
    }

    double *q(u64 j) { return Q.get() + j * ld; }

    double t_grow = 0.0; // SVDLAS2_TRACE: host time spent growing the panel
    void ensure_cols(u64 need) {
        if (need <= qcap) return;
        const double tg0 = wall();
        struct TG { double &acc; double t0; ~TG() { acc += wall() - t0; } } tg{t_grow, tg0};
        u64 cap = qcap ? qcap * 2 : std::max<u64>(64, 2 * (std::max<u64>(dims, 8) + dims));
        if (cap < need) cap = need;
        if (cap > iters + 1) cap = std::max(need, iters + 1);
        DevBuf<double> nq(cap * ld);
        if (qused) LAS2_CUDA(cudaMemcpyAsync(nq.get(), Q.get(), qused * ld * sizeof(double), cudaMemcpyDeviceToDevice, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
        Q = std::move(nq);
        qcap = cap;
    }

    void opb(const double *x, double *y) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (op.profile) {
            LAS2_CUDA(cudaEventCreate(&e0)); LAS2_CUDA(cudaEventCreate(&e1));
            LAS2_CUDA(cudaEventRecord(e0, s));
        }
        op.opb(x, y, t.get(), &prof.n_kernel_launches);
        if (op.profile) {
            LAS2_CUDA(cudaEventRecord(e1, s));
            opb_events.emplace_back(e0, e1);
        }
        prof.n_opb += 1;
        prof.n_spmv += 2;
    }

    void opb_step(const double *x, const double *qprev, double rnm, PartialSlot &alf) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (op.profile) {
            LAS2_CUDA(cudaEventCreate(&e0)); LAS2_CUDA(cudaEventCreate(&e1));
            LAS2_CUDA(cudaEventRecord(e0, s));
        }
        PartialSlot out = alf;
        op.opb_step(x, qprev, rnm, r.get(), t.get(), &out, ld, &prof.n_kernel_launches);
        alf = out; // (the ring hands out its own partial buffer)
        if (op.profile) {
            LAS2_CUDA(cudaEventRecord(e1, s));
            opb_events.emplace_back(e0, e1);
        }
        prof.n_opb += 1;
        prof.n_spmv += 2;
    }

    // read back the canonical sums of the given partial slots (one sync)
    void gather(std::initializer_list<int> which, double *out) {
        GatherList g;
        g.count = 0;
        for (int w : which) { g.parts[g.count] = slot[w].parts; g.nparts[g.count] = slot[w].nparts; g.count++; }
        gather_scalars(g, hs.get(), s);
        prof.n_kernel_launches += 1;
        const double tw0 = wall();
        LAS2_CUDA(cudaStreamSynchronize(s));
        prof.t_sync_wait += wall() - tw0;
        prof.n_host_syncs += 1;
        for (int i = 0; i < g.count; i++) out[i] = hs[i];
    }

    Coef sum_of(int w) const { return Coef::partials(slot[w].parts, slot[w].nparts); }

    // wait for the step chain's scalars: the kernel's last act is a sequence number in pinned memory (no stream sync)
    void wait_chain(double *out) {
        volatile unsigned long long *flag = reinterpret_cast<volatile unsigned long long *>(hs.get() + 4);
        const double tw0 = wall();
        static const bool poll = [] { const char *e = std::getenv("SVDLAS2_CHAIN_POLL"); return !(e && e[0] == '0'); }();
        if (!poll) LAS2_CUDA(cudaStreamSynchronize(s)); // experiment knob: plain stream sync instead of polling the flag
        u64 spins = 0;
        while (*flag != chain_seq) {
            if ((++spins & 0xfff) == 0) {
                const cudaError_t e = cudaStreamQuery(s);
                if (e != cudaSuccess && e != cudaErrorNotReady) LAS2_CUDA(e);
                if (e == cudaSuccess && *flag != chain_seq) { // finished without delivering: a launch failure
                    LAS2_CUDA(cudaGetLastError());
                    if (*flag != chain_seq) throw Error(SVDLAS2_ERR_CUDA, "step chain finished without its scalars");
                }
            }
        }
        prof.t_sync_wait += wall() - tw0;
        prof.n_host_syncs += 1;
        for (int i = 0; i < 4; i++) out[i] = hs[i];
    }

    // ---- startv (src/lib.rs:1098-1147): r <- B'B(random); on a restart also swept against q_0..q_{step-1} ----
    double startv(u64 step, uint32_t seed) {
        {
            std::vector<double> h(ld, 0.0);
            ChaCha12Stream g(seed); // rebuilt from the same seed on every call, :1109-1114
            if (op.relabel.active) { // the same vector, element c of the caller's numbering at its new position
                std::vector<double> ho(N);
                for (u64 i = 0; i < N; i++) ho[i] = g.next_unit_pm1();
                const std::vector<u32> &n2o = op.relabel.h_new2old;
                for (u64 i = 0; i < N; i++) h[i] = ho[n2o[i]];
            } else {
                for (u64 i = 0; i < N; i++) h[i] = g.next_unit_pm1();
            }
            LAS2_CUDA(cudaMemcpyAsync(xs.get(), h.data(), ld * sizeof(double), cudaMemcpyHostToDevice, s));
            LAS2_CUDA(cudaStreamSynchronize(s));
            prof.h2d_bytes += ld * sizeof(double);
        }
        opb(xs.get(), r.get()); // apply the operator to put r in range (essential if B singular)
        vec_dot(r.get(), r.get(), ld, slot[0], s);
        prof.n_kernel_launches += 1;
        double rnm2;
        gather({0}, &rnm2);
        // the reference retries up to three times with the SAME regenerated vector, so one try decides
        // (a NaN passes this test exactly as in the reference, :1127, and surfaces later as ImtqlbError)
        if (rnm2 <= 0.0) throw Error(SVDLAS2_ERR_STARTV, "rnm2 <= 0.0, rnm2 = " + std::to_string(rnm2));
        if (step > 0) {
            // w0 -= (w3 . q_i) q_i with w3 the frozen copy: classical GS against all stored vectors, :1132-1135
            panel_cgs2(Q.get(), ld, N, 0, (u32)step, r.get(), nullptr, pw, &slot[1], &slot[2], nullptr, s,
                       &prof.n_kernel_launches, nullptr, nullptr, nullptr, op.comm, &n_sharded_sweeps);
            // make sure q[step] is orthogonal to q[step-1], :1138
            vec_dot(q(step - 1), r.get(), ld, slot[3], s);
            vec_axpy_dot(r.get(), q(step - 1), sum_of(3), nullptr, ld, slot[4], s);
            prof.n_kernel_launches += 2;
            double dot;
            gather({4}, &dot);
            rnm2 = (dot <= kEps * rnm2) ? 0.0 : dot;
        }
        return std::sqrt(rnm2);
    }

    // ---- stpone (src/lib.rs:1179-1201) ----
    void stpone(uint32_t seed, double *rnm, double *tol) {
        double n0 = startv(0, seed);
        if (near(n0, 0.0)) throw Error(SVDLAS2_ERR_STPONE, "rnm == 0.0");
        ensure_cols(1);
        vec_scale(q(0), r.get(), 1.0 / n0, ld, s); // normalize starting vector
        qused = 1;
        opb(q(0), r.get());                         // take the first step
        vec_dot(r.get(), q(0), ld, slot[0], s);
        vec_axpy_dot(r.get(), q(0), sum_of(0), q(0), ld, slot[1], s);
        vec_axpy_dot(r.get(), q(0), sum_of(1), nullptr, ld, slot[2], s);
        prof.n_kernel_launches += 4;
        double v[3];
        gather({0, 1, 2}, v);
        alf[0] = v[0];
        alf[0] += v[1];
        *rnm = std::sqrt(v[2]);
        const double anorm = *rnm + std::fabs(alf[0]);
        *tol = std::sqrt(kEps) * anorm;
    }

    // ---- ortbnd (src/lib.rs:1431-1453): the eta recurrence, host O(j) ----
    void ortbnd(u64 step, double rnm) {
        if (step < 1) return;
        if (!near(rnm, 0.0) && step > 1) {
            oldeta[0] = (bet[1] * eta[1] + (alf[0] - alf[step]) * eta[0] - bet[step] * oldeta[0]) / rnm + eps1;
            for (u64 i = 1; i + 2 <= step; i++)
                oldeta[i] = (bet[i + 1] * eta[i + 1] + (alf[i] - alf[step]) * eta[i] + bet[i] * eta[i - 1] -
                             bet[step] * oldeta[i]) / rnm + eps1;
        }
        oldeta[step - 1] = eps1;
        eta.swap(oldeta);
        eta[step] = eps1;
    }

    // ---- purge (src/lib.rs:1356-1400): reorthogonalise r and q_step against q_ll..q_{step-1} when eta says so ----
    void purge(u64 ll, u64 step, double *rnm, double tol) {
        if (step < ll + 2) return;
        const double reps = std::sqrt(kEps);
        // NB the scan starts at eta[0] and ll is added afterwards (reference behaviour, :1364)
        const u64 k = argmax_abs(step - (ll + 1), eta.data()) + ll;
        if (!(std::fabs(eta[k]) > reps)) return;
        prof.n_purge_fired += 1;
        const double reps1 = eps1 / reps;
        bool again = true;
        for (int pass = 0; pass < 2 && again; pass++) {
            if (*rnm > tol) {
                // c = Q'[q r] against the frozen pair, [q r] -= Q c, then t = r . q (new), :1374-1384
                panel_cgs2(Q.get(), ld, N, (u32)ll, (u32)(step - ll), q(step), r.get(), pw, &slot[0], &slot[1],
                           &slot[2], s, &prof.n_kernel_launches, nullptr, nullptr, nullptr, op.comm, &n_sharded_sweeps);
                vec_axpy_dot(r.get(), q(step), sum_of(2), nullptr, ld, slot[3], s); // r -= t q ; |r|^2, :1386-1388
                prof.n_kernel_launches += 1;
                double v[4];
                gather({0, 1, 2, 3}, v);
                const double tq = v[0];
                const double tr = v[1] + std::fabs(v[2]);
                *rnm = std::sqrt(v[3]);
                prof.n_purge_passes += 1;
                prof.n_purge_vectors += step - ll;
                if (tq <= reps1 && tr <= *rnm * reps1) again = false;
            }
        }
        for (u64 i = ll; i <= step; i++) { eta[i] = eps1; oldeta[i] = eps1; }
    }

    // ---- lanczos_step (src/lib.rs:1233-1319) ----
    u64 lanczos_step(u64 first, u64 last, u64 *ll, bool *enough, double *rnm, double *tol) {
        u64 j = first;
        while (j < last) {
            // (the reference's swaps / storq / storp are implicit: q_{j-1} already sits in panel column j-1)
            bet[j] = *rnm;
            // restart if invariant subspace is found
            if (near(*rnm, 0.0)) {
                prof.n_restarts += 1;
                *rnm = startv(j, 0); // restarts always use seed 0, :1259
                if (near(*rnm, 0.0)) *enough = true;
            }
            if (*enough) break;

            // take a lanczos step
            ensure_cols(j + 1);
            const double rn = *rnm;
            if (spec_col != j) vec_scale(q(j), r.get(), 1.0 / rn, ld, s); // (else: the previous step's chain already wrote it)
            spec_col = ~(u64)0;
            qused = j + 1;
            int ns = 0; // slots in use for this step
            // r = B'B q_j - rnm q_{j-1} ; alf = r . q_j  (sharded: the all-reduce, the update and the dot are ONE kernel)
            // (slot 7 is reserved for alf: in a sharded run it points into the ring's partial buffer)
            constexpr int s_alf = 7;
            const bool chain_now = use_chain && j > kMaxLL;
            const bool fold_epilogue = chain_now && !op.ring; // single GPU: the update r -= rnm q_{j-1} becomes the chain's first phase
            const bool fused_now = fold_epilogue && use_fused; // svd_opb itself runs inside the chain's launch
            if (fused_now) { prof.n_opb += 1; prof.n_spmv += 2; }
            else if (fold_epilogue) opb(q(j), r.get());
            else opb_step(q(j), q(j - 1), rn, slot[s_alf]);
            if (j <= kMaxLL) {
                // |alf[j-1]| > 4 |alf[j]| decides ll only while j <= MAXLL (:1282): needs alf[j] now
                double a;
                gather({s_alf}, &a);
                if (std::fabs(alf[j - 1]) > 4.0 * std::fabs(a)) *ll = j;
            }
            double v[4];
            const u64 purges_before = prof.n_purge_fired;
            if (chain_now) {
                // ---- :1277-1304 as ONE launch (vec_step_chain), scalars through pinned memory ----
                const u64 lim = std::min<u64>(j - 1, *ll);
                static const bool spec_on = [] { const char *e = std::getenv("SVDLAS2_CHAIN_SPEC"); return !(e && e[0] == '0'); }();
                const bool speculate = spec_on && (j + 1 < last || j + 1 < spec_horizon);
                if (speculate) ensure_cols(j + 2);
                StepChain c;
                const int o = fold_epilogue ? 1 : 0; // phases shift by one when the epilogue leads the chain
                c.nphases = (int)lim + 3 + o;
                if (fold_epilogue) { c.x[0] = q(j - 1); c.y[0] = q(j); } // r -= rnm q_{j-1} ; alf = r . q_j
                c.x[o] = q(j);
                c.y[o] = lim > 0 ? q(0) : q(j - 1);
                for (u64 i = 1; i <= lim; i++) {
                    c.x[o + i] = q(i - 1);
                    c.y[o + i] = i < lim ? q(i) : q(j - 1);
                    eta[i - 1] = eps1;
                    oldeta[i - 1] = eps1;
                }
                c.x[o + lim + 1] = q(j - 1); c.y[o + lim + 1] = q(j);
                c.x[o + lim + 2] = q(j);     c.y[o + lim + 2] = nullptr;
                c.pick[0] = o; c.pick[1] = o + (int)lim + 1; c.pick[2] = o + (int)lim + 2; c.pick[3] = 0;
                c.alf_parts = fold_epilogue ? nullptr : slot[s_alf].parts;
                c.alf_nparts = fold_epilogue ? 0 : slot[s_alf].nparts;
                c.coef0 = rn;
                c.q_next = speculate ? q(j + 1) : nullptr;
                chain_seq += 1;
                if (fused_now) vec_step_fused(r.get(), ld, c, fused, chain_work.get(), hs.get(), chain_seq, s);
                else vec_step_chain(r.get(), ld, c, chain_work.get(), hs.get(), chain_seq, s);
                prof.n_kernel_launches += 1;
                wait_chain(v);
                if (speculate) spec_col = j + 1;
            } else {
            // r -= alf q_j, then orthogonalize against the first ll Lanczos vectors (:1285-1292) ...
            const double *ax = q(j);
            int prev = s_alf;
            const u64 lim = std::min<u64>(j - 1, *ll);
            for (u64 i = 0; i < lim; i++) {
                vec_axpy_dot(r.get(), ax, sum_of(prev), q(i), ld, slot[ns], s);
                prof.n_kernel_launches += 1;
                ax = q(i);
                prev = ns;
                ns++;
                eta[i] = eps1;
                oldeta[i] = eps1;
            }
            // ... and the extended local reorthogonalization (:1295-1304)
            vec_axpy_dot(r.get(), ax, sum_of(prev), q(j - 1), ld, slot[ns], s); // t1 = r . q_{j-1}
            const int s_t1 = ns++;
            vec_axpy_dot(r.get(), q(j - 1), sum_of(s_t1), q(j), ld, slot[ns], s); // r -= t1 q_{j-1} ; t2 = r . q_j
            const int s_t2 = ns++;
            vec_axpy_dot(r.get(), q(j), sum_of(s_t2), nullptr, ld, slot[ns], s); // r -= t2 q_j ; |r|^2
            const int s_nrm = ns++;
            prof.n_kernel_launches += 3;
            gather({s_alf, s_t1, s_t2, s_nrm}, v); // the one read-back of this step
            }
            const double nrm2 = v[3];
            alf[j] = v[0];
            if (bet[j] > 0.0) bet[j] += v[1];
            alf[j] += v[2];
            *rnm = std::sqrt(nrm2);
            const double anorm = bet[j] + std::fabs(alf[j]) + *rnm;
            *tol = std::sqrt(kEps) * anorm;

            ortbnd(j, *rnm);                 // update the orthogonality bounds
            const double tp0 = wall();
            purge(*ll, j, rnm, *tol);        // restore the orthogonality state when needed
            t_purge_host += wall() - tp0;
            if (prof.n_purge_fired != purges_before) spec_col = ~(u64)0; // r changed: the speculative q_{j+1} is stale
            if (*rnm <= *tol) { *rnm = 0.0; spec_col = ~(u64)0; }
            j += 1;
        }
        return j;
    }

    // ---- the analysis of T at the end of a batch (src/lib.rs:1975-2010): eigenvalues + last-row eigenvector components of
    // every unreduced block, error bounds, number of stabilised Ritz values.  Reads alf[0..j] and betv[0..j+1]; writes ritz /
    // bnd / w5 only, so it may run next to the Lanczos steps (lanso_overlapped). ----
    void analyze_T(u64 j, double rnm, double tol, const double *betv, bool *enough, u64 *neig) {
        const double ta0 = wall();
        struct TAcc { double &acc; double t0; ~TAcc() { acc += wall() - t0; } } tacc{t_analysis_host, ta0};
        u64 l = 0;
        for (u64 id2 = 0; id2 < j; id2++) {
            if (l > j) break;
            u64 i = l;
            while (i <= j && !near(betv[i + 1], 0.0)) i++;
            i = std::min(i, j);
            const u64 sz = i - l;
            for (u64 k = 0; k <= sz; k++) ritz[l + sz - k] = alf[l + k];     // svd_dcopy reverses, :1985
            for (u64 k = 0; k < sz; k++) w5[l + 1 + (sz - 1) - k] = betv[l + 1 + k]; // :1986
            // T in one unreduced block: this is the very iteration imtql2 will run on T if the analysis turns out to be the last
            // one, so its rotations are kept for ritvec
            bool whole = l == 0 && sz == j;
            if (whole) {
                // ~1.1 (j+1)^2 rotations; make room before recording (the vector is empty here: no copy), geometrically, and
                // never more than 1 GB ahead of need (beyond that the vector grows on demand, as ql_rotations' own would)
                const size_t need = std::min<size_t>((size_t)(2.6 * (double)(j + 1) * (double)(j + 1)) + 64, (size_t)1 << 27);
                kept_plan.sc.clear();
                try {
                    if (kept_plan.sc.capacity() < need) kept_plan.sc.reserve(std::max(need, 2 * kept_plan.sc.capacity()));
                } catch (const std::bad_alloc &) {
                    whole = false; // no room for a plan now: ritvec will run imtql2's recurrence itself
                }
            }
            int rc = ql_values(sz + 1, &ritz[l], &w5[l], &bnd[l], whole ? &kept_plan : nullptr);
            kept_plan_js = (whole && rc == 0) ? j + 1 : 0;
            if (rc == SVDLAS2_ERR_IMTQLB)
                throw Error(rc, "imtqlb no convergence to an eigenvalue after 30 iterations");
            if (rc) throw Error(rc, "imtqlb: expected 'm' to be non-zero");
            for (u64 m = l; m <= i; m++) bnd[m] = rnm * std::fabs(bnd[m]);
            l = i + 1;
        }
        cosort_ascending(j + 1, ritz.data(), bnd.data());
        size_t ne = 0;
        int rc = refine_bounds(enough, endl_, endr_, ritz.data(), bnd.data(), j, tol, &ne);
        if (rc) throw Error(rc, "error_bound: expected 'step' to be non-zero");
        *neig = ne;
    }
    double endl_ = 0.0, endr_ = 0.0;
    QlPlan kept_plan;      // rotation plan of the latest analysis (valid for T of order kept_plan_js; 0: none)
    u64 kept_plan_js = 0;

    // the batch scheduler (:1996-2009): returns the next `last`, or sets *enough
    u64 next_last(u64 first, u64 j, u64 neig, u64 *intro, bool *enough) const {
        u64 last = first;
        if (neig < dims) {
            if (neig == 0) {
                last = first + 9;
                *intro = first;
            } else {
                last = first + std::max<u64>(3, 1 + ((j - *intro) * (dims - neig)) / neig);
            }
            last = std::min(last, iters);
        } else {
            *enough = true;
        }
        *enough = *enough || first >= iters;
        return last;
    }

    // ---- lanso (src/lib.rs:1925-2017) ----
    u64 lanso(const double end_interval[2], uint32_t seed, u64 *neig_out) {
        endl_ = end_interval[0];
        endr_ = end_interval[1];
        double rnm, tol;
        stpone(seed, &rnm, &tol);
        eta[0] = eps1;
        oldeta[0] = eps1;
        // The Lanczos recurrence does not depend on the analyses (they only decide where it STOPS), so on one GPU they can run
        // on a helper thread while the steps go on (lanso_overlapped, SVDLAS2_OVERLAP_ANALYSIS=1).  Opt-in: the cost of an
        // analysis grows with j^2, so the LAST one dominates their sum and nothing useful can run next to it -- measured
        // (profiles/r2y_overlapped_analysis.md): cfg 1 unchanged, cfg 5 slower (114 discarded steps with their purges, and two
        // busy host threads inside the box's CPU quota).  Sharded runs always keep the sequential schedule: every rank must
        // launch the same collectives, and how far a rank runs ahead of its helper depends on timing.
        const char *e = std::getenv("SVDLAS2_OVERLAP_ANALYSIS");
        if (!op.comm && e && e[0] == '1') return lanso_overlapped(rnm, tol, neig_out);
        u64 ll = 0, first = 1;
        u64 last = std::min(iters, std::max<u64>(dims, 8) + dims);
        bool enough = false;
        u64 j = 0, intro = 0, neig = 0;

        while (!enough) {
            if (rnm <= tol) rnm = 0.0;
            const u64 steps = lanczos_step(first, last, &ll, &enough, &rnm, &tol);
            j = enough ? steps - 1 : last - 1;
            first = j + 1;
            bet[first] = rnm;
            analyze_T(j, rnm, tol, bet.data(), &enough, &neig);
            // should we stop?
            last = next_last(first, j, neig, &intro, &enough);
        }
        *neig_out = neig;
        return j;
    }

    // lanso with the analyses of T overlapped (single GPU).  The main thread runs the steps one after the other and records,
    // per finished step, the (rnm, tol) the sequential code would hand to an analysis there; the helper thread walks the
    // batch boundaries exactly as the loop above does (same first / last / intro arithmetic), analysing T at a boundary as soon
    // as that step has been taken, and raises `stop` when the reference would leave the loop.  Steps the main thread took
    // beyond that boundary are discarded: they only ever wrote panel columns, alf / bet entries and r BEYOND the boundary
    // (purge touches the current column and r), and the counters are put back to their values at the boundary.  Errors keep
    // the reference's order: an exception of step s surfaces only if no earlier boundary stops or fails first.
    struct Overlap {
        std::mutex mu;
        std::condition_variable cv;
        // main -> helper
        u64 done_step = 0;          // steps 1 .. done_step are complete
        bool halted = false;        // the main thread takes no more steps: breakdown (`enough` inside a step), or an exception
        u64 halt_first = 0;         // ... the step that was not taken
        bool halt_is_error = false;
        double halt_rnm = 0.0, halt_tol = 0.0;
        std::exception_ptr main_exc;
        u64 awaited = ~(u64)0;      // boundary step the helper is waiting for right now
        // helper -> main
        bool decided = false;
        u64 j_final = 0, neig_final = 0;
        std::exception_ptr helper_exc;
        std::atomic<bool> stop{false};
    };

    u64 lanso_overlapped(double rnm, double tol, u64 *neig_out) {
        Overlap ov;
        std::vector<double> rnm_at(iters + 1, 0.0), tol_at(iters + 1, 0.0);
        std::vector<svdlas2_profile> prof_at(iters + 1);
        rnm_at[0] = rnm; tol_at[0] = tol; prof_at[0] = prof;

        std::thread helper([&] {
            try {
                std::vector<double> betv(iters + 2, 0.0);
                u64 first = 1, last = std::min(iters, std::max<u64>(dims, 8) + dims);
                u64 j = 0, intro = 0, neig = 0;
                bool enough = false;
                while (!enough) {
                    double a_rnm, a_tol;
                    {
                        std::unique_lock<std::mutex> lk(ov.mu);
                        ov.awaited = last - 1;
                        ov.cv.notify_all();
                        ov.cv.wait(lk, [&] { return ov.done_step + 1 >= last || ov.halted; });
                        ov.awaited = ~(u64)0;
                        if (ov.done_step + 1 >= last) {
                            j = last - 1;                       // the batch ran to its end
                            a_rnm = rnm_at[j]; a_tol = tol_at[j];
                        } else {
                            if (ov.halt_is_error) std::rethrow_exception(ov.main_exc); // the batch would have thrown at that step
                            enough = true;                      // lanczos_step left the batch early (:1262): j = steps - 1
                            j = ov.halt_first - 1;
                            a_rnm = ov.halt_rnm; a_tol = ov.halt_tol;
                        }
                    }
                    first = j + 1;
                    std::memcpy(betv.data(), bet.data(), (j + 1) * sizeof(double)); // final: steps <= j are complete
                    betv[first] = a_rnm;                                              // bet[first] = rnm, :1971
                    analyze_T(j, a_rnm, a_tol, betv.data(), &enough, &neig);
                    last = next_last(first, j, neig, &intro, &enough);
                }
                std::lock_guard<std::mutex> lk(ov.mu);
                ov.j_final = j; ov.neig_final = neig; ov.decided = true;
            } catch (...) {
                std::lock_guard<std::mutex> lk(ov.mu);
                ov.helper_exc = std::current_exception();
                ov.decided = true;
            }
            ov.stop.store(true, std::memory_order_release);
            ov.cv.notify_all();
        });
        // leaving by an exception that is not a step's own (allocation failure ...): release the helper before joining it
        struct Joiner {
            std::thread &t; Overlap &ov;
            ~Joiner() {
                if (!t.joinable()) return;
                {
                    std::lock_guard<std::mutex> lk(ov.mu);
                    if (!ov.decided && !ov.halted) {
                        ov.halted = true; ov.halt_is_error = true; ov.halt_first = ov.done_step + 1;
                        ov.main_exc = std::make_exception_ptr(Error(SVDLAS2_ERR_CUDA, "solve abandoned"));
                    }
                }
                ov.cv.notify_all();
                t.join();
            }
        } joiner{helper, ov};

        u64 ll = 0, j = 1;
        bool enough = false;
        spec_horizon = iters;
        double step_s = 0.0; // running mean of a step's wall time
        u64 nsteps = 0;
        auto halt = [&](bool is_error) {
            std::unique_lock<std::mutex> lk(ov.mu);
            ov.halted = true; ov.halt_first = j; ov.halt_is_error = is_error; ov.halt_rnm = rnm; ov.halt_tol = tol;
            if (is_error) ov.main_exc = std::current_exception();
            ov.cv.notify_all();
            ov.cv.wait(lk, [&] { return ov.decided; });
        };
        while (!ov.stop.load(std::memory_order_acquire)) {
            if (j >= iters) { // every step a batch can ask for has been taken: the helper decides alone from here
                std::unique_lock<std::mutex> lk(ov.mu);
                ov.cv.wait(lk, [&] { return ov.decided; });
                break;
            }
            const double ts0 = wall();
            try {
                if (rnm <= tol) rnm = 0.0;
                lanczos_step(j, j + 1, &ll, &enough, &rnm, &tol);
            } catch (...) {
                halt(true);
                break;
            }
            if (enough) { halt(false); break; } // breakdown: the step was not taken
            step_s += (wall() - ts0 - step_s) / (double)(++nsteps);
            rnm_at[j] = rnm; tol_at[j] = tol; prof_at[j] = prof;
            bool at_boundary;
            {
                std::lock_guard<std::mutex> lk(ov.mu);
                ov.done_step = j;
                at_boundary = ov.awaited == j;
            }
            ov.cv.notify_all();
            // A step that is long next to an analysis is not worth gambling: give the helper a third of a step's time to
            // answer before the next step is launched (it usually says "stop" or "go on" within that).
            if (at_boundary && step_s > 50e-6) {
                std::unique_lock<std::mutex> lk(ov.mu);
                ov.cv.wait_for(lk, std::chrono::duration<double>(0.3 * step_s),
                               [&] { return ov.decided || (ov.awaited != ~(u64)0 && ov.awaited != j); });
            }
            j++;
        }
        helper.join();
        if (ov.helper_exc) std::rethrow_exception(ov.helper_exc);
        // the state the sequential loop would have ended with
        const u64 jf = ov.j_final;
        const double t_sync = prof.t_sync_wait;
        const u64 launches = prof.n_kernel_launches, syncs = prof.n_host_syncs, h2d = prof.h2d_bytes;
        if (!ov.halted || jf < ov.halt_first - 1 || ov.halt_is_error) {
            prof = prof_at[jf];
            bet[jf + 1] = rnm_at[jf];
        } else {
            bet[jf + 1] = ov.halt_rnm; // the batch that broke down: counters are already those of its last step
        }
        prof.t_sync_wait = t_sync; prof.n_kernel_launches = launches; prof.n_host_syncs = syncs; prof.h2d_bytes = h2d; // work really done
        n_overshoot = ov.done_step > jf ? ov.done_step - jf : 0;
        spec_col = ~(u64)0;
        qused = std::min(qused, jf + 1);
        *neig_out = ov.neig_final;
        return jf;
    }
    u64 n_overshoot = 0;  // steps taken beyond the final boundary (lanso_overlapped)
    u64 spec_horizon = 0; // lanso_overlapped calls lanczos_step one step at a time: the next step exists up to here

    // ---- ritvec (src/lib.rs:1767-1879) ----
    void ritvec(u64 steps, u64 neig, double kappa, svdlas2_result *res, double *t_imtql2) {
        const u64 js = steps + 1;
        std::vector<double> dd(js), ee(js + 1, 0.0);
        for (u64 i = 0; i < js; i++) dd[js - 1 - i] = alf[i];                   // svd_dcopy(js, 0, alf, ..), :1790
        for (u64 i = 0; i < steps; i++) ee[1 + (steps - 1) - i] = bet[1 + i];  // svd_dcopy(steps, 1, bet, w5), :1791

        const double tq0 = wall();
        QlPlan own_plan;
        static const bool reuse_on = [] { const char *e = std::getenv("SVDLAS2_REUSE_QL_PLAN"); return !(e && e[0] == '0'); }();
        const bool reused = reuse_on && kept_plan_js == js;
        // (reused: the last analysis ran this very iteration on this very T, host_math.hpp; the plan stays where it is so
        // that its storage serves the next solve)
        const QlPlan &plan = reused ? kept_plan : own_plan;
        if (!reused) {
            int rc = ql_rotations(js, dd.data(), ee.data(), own_plan);
            if (rc == SVDLAS2_ERR_IMTQL2) throw Error(rc, "imtql2 no convergence to an eigenvalue after 30 iterations");
            if (rc) throw Error(rc, "imtql2: expected 'm' to be non-zero");
        }
        const double tq1 = wall();
        const u32 ldz = (u32)round_up(js, 32);
        DevBuf<double> zT((size_t)ldz * js);
        ql_replay(plan, (u32)js, ldz, zT.get(), s, &prof.n_kernel_launches, &prof.h2d_bytes);
        LAS2_CUDA(cudaStreamSynchronize(s));
        *t_imtql2 = wall() - tq0;
        if (std::getenv("SVDLAS2_TRACE"))
            fprintf(stderr, "[svdlas2] imtql2: js=%llu rotations=%llu sweeps=%zu host %.4f s%s, device replay %.4f s\n",
                    (unsigned long long)js, (unsigned long long)plan.rotations(), plan.sweeps.size(), tq1 - tq0,
                    reused ? " (plan of the last analysis reused)" : "", wall() - tq1);

        // acceptance (:1800-1801) in ascending eigenvalue order; the reference's circular buffer + rotate_array
        // keep the LAST `dims` accepted, largest first (:1802-1827)
        std::vector<u64> accepted;
        for (u64 k = 0; k < js; k++)
            if (bnd[k] <= kappa * std::fabs(ritz[k]) && k + 1 > js - neig) accepted.push_back(k);
        const u64 nsig = accepted.size();
        const u64 d = std::min(dims, nsig);
        res->d = d;
        res->diagnostics.significant_values = d;   // sic, :648
        res->diagnostics.singular_values = nsig;   // sic, :649

        const u64 mcols = M;
        // the N-side vectors are replicated over the ranks of a sharded run: on request only rank 0 downloads them
        // (the other ranks return an N-side array of zero columns)
        u64 vb = 0, ve = 0;
        op.nside_rows(d, &vb, &ve);
        const u64 vrows = ve - vb;          // N-side vectors this rank brings to the host
        const u64 ncols = vrows ? N : 0;
        res->nside_row_begin = vb;
        res->nside_row_count = vrows;
        const bool trace = std::getenv("SVDLAS2_TRACE") != nullptr;
        const double ta0 = wall();
        double *hU = (double *)(prefetched ? fut_m.get() : host_acquire(std::max<u64>(1, d * mcols) * sizeof(double)));
        double *hV = nullptr;
        try {
            hV = (double *)(prefetched ? fut_n.get() : host_acquire(std::max<u64>(1, vrows * ncols) * sizeof(double)));
        } catch (...) { host_release(hU); throw; }
        double *hS = (double *)std::malloc(std::max<size_t>(1, d) * sizeof(double));
        if (!hS) { host_release(hU); host_release(hV); throw Error(SVDLAS2_ERR_ALLOC, "host result allocation failed"); }
        // hand ownership to the result right away so that an exception below cannot leak
        const bool tr = op.transposed;
        res->ut = tr ? hV : hU; res->ut_cols = tr ? ncols : mcols;
        res->vt = tr ? hU : hV; res->vt_cols = tr ? mcols : ncols;
        res->s = hS;
        if (d == 0) return;
        if (trace) fprintf(stderr, "[svdlas2] ritvec: host result buffers (%.2f GB) ready after %.3f s (prefetched=%d)\n",
                           (double)(d * (mcols + ncols) * 8) / 1e9, wall() - ta0, (int)prefetched);
        const double tc0 = wall();

        std::vector<u32> phys(d);
        for (u64 i = 0; i < d; i++) phys[i] = plan.column_of[accepted[nsig - 1 - i]];
        DevBuf<u32> dphys(d);
        LAS2_CUDA(cudaMemcpyAsync(dphys.get(), phys.data(), d * sizeof(u32), cudaMemcpyHostToDevice, s));
        DevBuf<double> C((size_t)js * d);
        ritz_coefficients(zT.get(), ldz, (u32)js, dphys.get(), (u32)d, C.get(), s, &prof.n_kernel_launches);
        DevBuf<double> V((size_t)d * ld);
        bool ritz_sharded = false;
        ritz_contract(Q.get(), ld, N, (u32)js, C.get(), (u32)d, V.get(), s, &prof.n_kernel_launches, op.comm, &ritz_sharded);
        LAS2_CUDA(cudaStreamSynchronize(s));
        if (trace) fprintf(stderr, "[svdlas2] ritvec: contraction N=%llu js=%llu d=%llu in %.3f s (%.1f TFLOP/s)%s\n",
                           (unsigned long long)N, (unsigned long long)js, (unsigned long long)d, wall() - tc0,
                           2.0 * (double)N * (double)js * (double)d / (wall() - tc0) / 1e12,
                           ritz_sharded ? ", rows split over the ranks" : "");
        const double tt0 = wall();
        zT.release();
        C.release();
        Q.release(); // the panel is no longer needed; make room for U
        qcap = qused = 0;

        // sigma_i = sqrt(v_i . B'B v_i), u_i = B v_i / sigma_i (:1839-1859).  The reference applies the operator
        // three times per triplet (B v and B'(B v) inside svd_opb for sigma, then B v again for u).  Here B v is
        // computed ONCE: v . B'(B v) == (B v) . (B v), so sigma = |B v| and u = B v / sigma come from the same
        // vector (a rounding-level difference in sigma, far inside the 1e-10 parity tolerance).  In a sharded run
        // every rank holds its row slice of B v: the squared norms are completed by a one-double all-reduce.
        // Finished rows stream to the (pinned) result on a second stream while later rows are still computed.
        cudaEvent_t ev;
        LAS2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        struct EvDel { cudaEvent_t e; ~EvDel() { cudaEventDestroy(e); } } evdel{ev};
        LAS2_CUDA(cudaEventRecord(ev, s));
        LAS2_CUDA(cudaStreamWaitEvent(copy_stream, ev, 0));
        DevBuf<double> Vout; // the N-side vectors in the caller's numbering (relabelled operators only)
        if (ncols && op.relabel.active) {
            Vout.alloc((size_t)vrows * N);
            unpermute_rows(V.get() + vb * ld, ld, N, (u32)vrows, op.relabel.old2new.get(), Vout.get(), s);
            prof.n_kernel_launches += 1;
            LAS2_CUDA(cudaEventRecord(ev, s));
            LAS2_CUDA(cudaStreamWaitEvent(copy_stream, ev, 0));
            LAS2_CUDA(cudaMemcpyAsync(hV, Vout.get(), vrows * N * sizeof(double), cudaMemcpyDeviceToHost, copy_stream));
        } else if (ncols) {
            LAS2_CUDA(cudaMemcpy2DAsync(hV, ncols * sizeof(double), V.get() + vb * ld, ld * sizeof(double), ncols * sizeof(double), vrows,
                                        cudaMemcpyDeviceToHost, copy_stream));
        }
        DevBuf<double> U((size_t)d * mcols);
        DevBuf<double> S(d);
        DevBuf<double> nrm2(d); // |B_g v_i|^2 per triplet (summed over ranks when sharded)
        // Matrices on the chunk stream take FOUR triplets per pass over the matrix (k_spmm4_chunk: one 32-byte gather and
        // one read of the 12 B/nnz stream serve four vectors; every column is bit-identical to its SpMV)
        static const bool no_spmm = std::getenv("SVDLAS2_NO_SPMM") != nullptr;
        const bool use4 = spmm4_supported(op.B) && !no_spmm;
        Spmm4Work w4;
        if (!op.comm && use4) {
            for (u64 i0 = 0; i0 < d; i0 += 4) {
                const u32 nv = (u32)std::min<u64>(4, d - i0);
                const int np = spmm4(op.B, V.get() + i0 * ld, ld, nv, U.get() + i0 * mcols, mcols, w4, s, &prof.n_kernel_launches);
                prof.n_spmv += nv;
                for (u32 r = 0; r < nv; r++)
                    vec_sigma_scale(U.get() + (i0 + r) * mcols, mcols, PartialSlot{w4.parts.get() + (size_t)r * kMaxParts, np},
                                    S.get() + i0 + r, s);
                prof.n_kernel_launches += nv;
                LAS2_CUDA(cudaEventRecord(ev, s));
                LAS2_CUDA(cudaStreamWaitEvent(copy_stream, ev, 0));
                LAS2_CUDA(cudaMemcpyAsync(hU + i0 * mcols, U.get() + i0 * mcols, nv * mcols * sizeof(double), cudaMemcpyDeviceToHost,
                                          copy_stream));
            }
        } else if (!op.comm) {
            for (u64 i = 0; i < d; i++) {
                double *vi = V.get() + i * ld;
                double *ui = U.get() + i * mcols;
                spmv(op.B, vi, ui, s, &prof.n_kernel_launches);
                prof.n_spmv += 1;
                vec_sumsq_any(ui, mcols, slot[0], s);
                vec_sigma_scale(ui, mcols, slot[0], S.get() + i, s);
                prof.n_kernel_launches += 2;
                LAS2_CUDA(cudaEventRecord(ev, s));
                LAS2_CUDA(cudaStreamWaitEvent(copy_stream, ev, 0));
                LAS2_CUDA(cudaMemcpyAsync(hU + i * mcols, ui, mcols * sizeof(double), cudaMemcpyDeviceToHost, copy_stream));
            }
        } else {
            // sharded: every rank holds its row slice of B v_i; the squared norms of a batch of triplets are completed
            // by ONE all-reduce (instead of one latency-bound collective per triplet), then the batch is scaled and
            // leaves for the host as one contiguous copy
            constexpr u64 kBatch = 32;
            for (u64 i0 = 0; i0 < d; i0 += kBatch) {
                const u64 nb = std::min(kBatch, d - i0);
                for (u64 i = i0; use4 && i < i0 + nb; i += 4) {
                    const u32 nv = (u32)std::min<u64>(4, i0 + nb - i);
                    const int np = spmm4(op.B, V.get() + i * ld, ld, nv, U.get() + i * mcols, mcols, w4, s, &prof.n_kernel_launches);
                    prof.n_spmv += nv;
                    GatherList g;
                    g.count = (int)nv;
                    for (u32 r = 0; r < nv; r++) { g.parts[r] = w4.parts.get() + (size_t)r * kMaxParts; g.nparts[r] = np; }
                    gather_scalars(g, nrm2.get() + i, s);
                    prof.n_kernel_launches += 1;
                }
                for (u64 i = i0; !use4 && i < i0 + nb; i++) {
                    double *ui = U.get() + i * mcols;
                    spmv(op.B, V.get() + i * ld, ui, s, &prof.n_kernel_launches);
                    prof.n_spmv += 1;
                    vec_sumsq_any(ui, mcols, slot[0], s);
                    GatherList g;
                    g.count = 1; g.parts[0] = slot[0].parts; g.nparts[0] = slot[0].nparts;
                    gather_scalars(g, nrm2.get() + i, s);
                    prof.n_kernel_launches += 2;
                }
                op.comm->allreduce_sum(nrm2.get() + i0, nb, s);
                for (u64 i = i0; i < i0 + nb; i++) {
                    vec_sigma_scale(U.get() + i * mcols, mcols, PartialSlot{nrm2.get() + i, 1}, S.get() + i, s);
                    prof.n_kernel_launches += 1;
                }
                LAS2_CUDA(cudaEventRecord(ev, s));
                LAS2_CUDA(cudaStreamWaitEvent(copy_stream, ev, 0));
                LAS2_CUDA(cudaMemcpyAsync(hU + i0 * mcols, U.get() + i0 * mcols, nb * mcols * sizeof(double),
                                          cudaMemcpyDeviceToHost, copy_stream));
            }
        }
        if (trace) {
            LAS2_CUDA(cudaStreamSynchronize(s));
            fprintf(stderr, "[svdlas2] ritvec: tail kernels done after %.3f s\n", wall() - tt0);
        }
        LAS2_CUDA(cudaMemcpyAsync(hS, S.get(), d * sizeof(double), cudaMemcpyDeviceToHost, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
        if (trace) fprintf(stderr, "[svdlas2] ritvec: sigma/U tail (%llu SpMV) in %.3f s\n", (unsigned long long)d, wall() - tt0);
        const double td0 = wall();
        LAS2_CUDA(cudaStreamSynchronize(copy_stream));
        if (trace) fprintf(stderr, "[svdlas2] ritvec: waited %.3f s more for the D2H stream\n", wall() - td0);
        prof.n_host_syncs += 2;
        prof.d2h_bytes += (vrows * ncols + d * mcols + d) * sizeof(double);
        prof.t_download = wall() - td0; // D2H time NOT hidden behind the sigma/U kernels
        const double tf0 = wall();
        U.release(); V.release(); S.release(); nrm2.release(); Vout.release();
        if (trace) fprintf(stderr, "[svdlas2] ritvec: device buffers released in %.3f s\n", wall() - tf0);
    }
};

} // namespace

Operator::~Operator() {
    if (stream) cudaStreamDestroy(stream);
}

void Operator::opb(const double *x, double *y, double *t, u64 *launches) {
    spmv(B, x, t, stream, launches);
    if (ring) {
        spmv(BT, t, ring->y_mine(), stream, launches);
        ring->reduce((N + 31) / 32 * 32, nullptr, 0.0, nullptr, nullptr, stream);
        if (launches) *launches += 1;
        if (y != ring->r_mine())
            LAS2_CUDA(cudaMemcpyAsync(y, ring->r_mine(), N * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        return;
    }
    spmv(BT, t, y, stream, launches);
    if (comm) comm->allreduce_sum(y, N, stream);
}

void Operator::opb_step(const double *x, const double *qprev, double rnm, double *r, double *t, PartialSlot *alf, u64 ld,
                        u64 *launches) {
    if (ring) {
        if (r != ring->r_mine()) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "opb_step: r must be the ring's buffer");
        spmv(B, x, t, stream, launches);
        spmv(BT, t, ring->y_mine(), stream, launches);
        ring->reduce(ld, qprev, rnm, x, alf, stream); // sum over ranks + r -= rnm qprev + partials of r . x, one kernel
        if (launches) *launches += 1;
        return;
    }
    opb(x, r, t, launches);
    if (qprev) {
        vec_axpy_dot(r, qprev, Coef::value(rnm), x, ld, *alf, stream);
        if (launches) *launches += 1;
    }
}

void plan_cache_trim() {
    PlanSpare &ps = plan_spare();
    std::lock_guard<std::mutex> lk(ps.mu);
    ps.plan = QlPlan();
}

void solve(Operator &op, u64 dimensions, u64 iterations, const double end_interval[2], double kappa,
           uint32_t random_seed, svdlas2_result *res) {
    const double t0 = wall();
    LAS2_CUDA(cudaSetDevice(op.device));
    if (!(random_seed > 0)) {
        random_seed = entropy_seed_u32(); // :578-581
        if (op.comm) {
            // every rank must build the SAME start vector: rank 0's draw wins (a sum in which the other ranks add zero;
            // a u32 is exact in a double)
            DevBuf<double> sd(1);
            const double mine = op.comm->rank() == 0 ? (double)random_seed : 0.0;
            LAS2_CUDA(cudaMemcpyAsync(sd.get(), &mine, sizeof(double), cudaMemcpyHostToDevice, op.stream));
            op.comm->allreduce_sum(sd.get(), 1, op.stream);
            double got = 0.0;
            LAS2_CUDA(cudaMemcpyAsync(&got, sd.get(), sizeof(double), cudaMemcpyDeviceToHost, op.stream));
            LAS2_CUDA(cudaStreamSynchronize(op.stream));
            random_seed = (uint32_t)got;
        }
    }

    // clamping uses the un-oriented shape (:583-594)
    const u64 mn = std::min(op.a_rows, op.a_cols);
    if (dimensions == 0 || dimensions > mn) dimensions = mn;
    if (iterations == 0 || iterations > mn) iterations = mn;
    else if (iterations < dimensions) iterations = dimensions;

    svdlas2_diagnostics &dg = res->diagnostics;
    dg.non_zero = op.nnz_global;
    dg.dimensions = dimensions;
    dg.iterations = iterations;
    dg.transposed = op.transposed ? 1 : 0;
    dg.end_interval[0] = end_interval[0];
    dg.end_interval[1] = end_interval[1];
    dg.random_seed = random_seed;
    dg.kappa = kappa;
    if (dimensions < 2)
        throw Error(SVDLAS2_ERR_LAS2, "svdLAS2: insufficient dimensions: " + std::to_string(dimensions)); // :596-600

    cudaEvent_t ev_begin, ev_end;
    LAS2_CUDA(cudaEventCreate(&ev_begin));
    LAS2_CUDA(cudaEventCreate(&ev_end));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } ev_guard{ev_begin, ev_end};
    LAS2_CUDA(cudaEventRecord(ev_begin, op.stream));

    if (std::getenv("SVDLAS2_TRACE")) { double z[6]; device_alloc_stats(z, true); }
    Solver sv(op, dimensions, iterations);
    u64 neig = 0;
    const double tl0 = wall();
    const u64 steps = sv.lanso(end_interval, random_seed, &neig);
    sv.prof.t_lanczos = wall() - tl0;
    if (std::getenv("SVDLAS2_TRACE") && op.comm)
        fprintf(stderr, "[svdlas2] lanso: %llu purge passes, %llu reorthogonalisation sweeps split over %d ranks\n",
                (unsigned long long)sv.prof.n_purge_passes, (unsigned long long)sv.n_sharded_sweeps, op.comm->world());
    if (std::getenv("SVDLAS2_TRACE"))
        fprintf(stderr, "[svdlas2] lanso: %.4f s, of which purge %.4f s (%llu passes, %llu stored vectors), analyses of T %.4f s%s, waiting "
                        "for step scalars %.4f s, growing the panel %.4f s; %llu steps taken beyond the final boundary\n", sv.prof.t_lanczos, sv.t_purge_host,
                (unsigned long long)sv.prof.n_purge_passes, (unsigned long long)sv.prof.n_purge_vectors, sv.t_analysis_host,
                sv.spec_horizon ? " (on the helper thread, next to the steps)" : "", sv.prof.t_sync_wait, sv.t_grow, (unsigned long long)sv.n_overshoot);

    kappa = std::max(std::fabs(kappa), eps34()); // :627
    dg.kappa = kappa;
    dg.lanczos_steps = steps + 1;
    dg.ritz_values_stabilized = neig;

    // trace for step-by-step parity tests
    const u64 js = steps + 1;
    res->trace_len = js;
    res->alf = (double *)std::malloc(js * sizeof(double));
    res->bet = (double *)std::malloc((js + 1) * sizeof(double));
    res->ritz = (double *)std::malloc(js * sizeof(double));
    res->bnd = (double *)std::malloc(js * sizeof(double));
    if (!res->alf || !res->bet || !res->ritz || !res->bnd) throw Error(SVDLAS2_ERR_ALLOC, "host trace allocation failed");
    std::memcpy(res->alf, sv.alf.data(), js * sizeof(double));
    std::memcpy(res->bet, sv.bet.data(), (js + 1) * sizeof(double));
    std::memcpy(res->ritz, sv.ritz.data(), js * sizeof(double));
    std::memcpy(res->bnd, sv.bnd.data(), js * sizeof(double));

    const double tr0 = wall();
    double t_imtql2 = 0.0;
    sv.ritvec(steps, neig, kappa, res, &t_imtql2);
    sv.prof.t_ritvec = wall() - tr0;
    sv.prof.t_imtql2 = t_imtql2;

    if (op.ring) ring_quiesce(*op.ring, op.stream); // no rank touches a peer's memory after its solve has returned
    LAS2_CUDA(cudaEventRecord(ev_end, op.stream));
    LAS2_CUDA(cudaEventSynchronize(ev_end));
    {
        float ms = 0.f;
        LAS2_CUDA(cudaEventElapsedTime(&ms, ev_begin, ev_end));
        sv.prof.t_device = ms * 1e-3;
    }
    if (op.profile) {
        double ms_total = 0.0;
        for (auto &e : sv.opb_events) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e.first, e.second) == cudaSuccess) ms_total += ms;
        }
        sv.prof.t_opb_device = ms_total * 1e-3;
    }
    sv.prof.t_upload = op.t_upload;
    sv.prof.h2d_bytes += op.upload.h2d_bytes;
    sv.prof.opb_algorithmic_bytes = op.opb_bytes();
    sv.prof.t_total = wall() - t0;
    if (std::getenv("SVDLAS2_TRACE")) {
        double a[6];
        device_alloc_stats(a, false);
        fprintf(stderr, "[svdlas2] solve: %.4f s; device allocator: %.0f allocations %.4f s (longest %.4f s), %.0f releases %.4f s (longest %.4f s)\n",
                sv.prof.t_total, a[2], a[0], a[1], a[5], a[3], a[4]);
    }
    res->profile = sv.prof;
}

} // namespace las2
