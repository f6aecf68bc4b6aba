// abi.cu -- the extern "C" boundary declared in include/svdlas2.h.  Exceptions never cross it: every entry
// point maps las2::Error (and anything else) to a status code and a thread-local message.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "host_pool.hpp"
#include "las2.hpp"
#include "../../include/svdlas2_testing.h"
#include "spmv.cuh"
#include "tridiag_device.cuh"

using namespace las2;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

template <typename F>
int guarded(F &&f) {
    try {
        g_last_error.clear();
        f();
        return SVDLAS2_OK;
    } catch (const Error &e) {
        cudaGetLastError(); // clear a sticky-free error state
        return fail(e.code, e.what());
    } catch (const std::bad_alloc &) {
        return fail(SVDLAS2_ERR_ALLOC, "host allocation failed");
    } catch (const std::exception &e) {
        return fail(SVDLAS2_ERR_CUDA, std::string("unexpected: ") + e.what());
    } catch (...) {
        return fail(SVDLAS2_ERR_CUDA, "unexpected exception");
    }
}

double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void select_device(int device, Operator &op) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        throw Error(SVDLAS2_ERR_CUDA, std::string("no usable CUDA device (this library has no CPU fallback): ") +
                                          cudaGetErrorString(e));
    }
    if (device < 0) LAS2_CUDA(cudaGetDevice(&device));
    if (device >= count) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "device index out of range");
    LAS2_CUDA(cudaSetDevice(device));
    op.device = device;
    LAS2_CUDA(cudaStreamCreateWithFlags(&op.stream, cudaStreamNonBlocking));
}

Operator *create_operator(int format, u64 nrows, u64 ncols, u64 nnz, const u64 *a, const u64 *b, const double *values,
                          int device) {
    if (format != SVDLAS2_FMT_CSR && format != SVDLAS2_FMT_CSC && format != SVDLAS2_FMT_COO)
        throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "unknown sparse format");
    std::unique_ptr<Operator> op(new Operator());
    select_device(device, *op);
    const double t0 = wall_now();
    op->a_rows = nrows; op->a_cols = ncols; op->nnz_global = nnz;
    // If the matrix is wide, the SVD is computed on its transpose (src/lib.rs:606)
    op->transposed = (double)ncols >= ((double)nrows * 1.2);
    op->M_global = op->M = op->transposed ? ncols : nrows;
    op->N = op->transposed ? nrows : ncols;
    HostSparse A{format, nrows, ncols, nnz, a, b, values};
    build_operator(A, op->transposed, op->stream, op->B, op->BT, op->upload, op->relabel);
    op->t_upload = wall_now() - t0;
    if (std::getenv("SVDLAS2_TRACE")) {
        double a[6];
        device_alloc_stats(a, true);
        fprintf(stderr, "[svdlas2] upload: %.4f s; device allocator: %.0f allocations %.4f s (longest %.4f s), %.0f releases %.4f s (longest %.4f s)\n",
                op->t_upload, a[2], a[0], a[1], a[5], a[3], a[4]);
    }
    return op.release();
}

svdlas2_result *new_result() {
    svdlas2_result *r = (svdlas2_result *)std::calloc(1, sizeof(svdlas2_result));
    if (!r) throw std::bad_alloc();
    return r;
}

int one_shot(int format, u64 nrows, u64 ncols, u64 nnz, const u64 *a, const u64 *b, const double *values,
             u64 dimensions, u64 iterations, const double end_interval[2], double kappa, uint32_t seed,
             svdlas2_result **out) {
    if (!out) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (!end_interval) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "end_interval is null");
    svdlas2_result *res = nullptr;
    int rc = guarded([&] {
        const double t0 = wall_now();
        res = new_result();
        // the reference rejects dimensions < 2 before touching the matrix (src/lib.rs:583-600)
        const u64 mn = nrows < ncols ? nrows : ncols;
        u64 dims = (dimensions == 0 || dimensions > mn) ? mn : dimensions;
        if (dims < 2) throw Error(SVDLAS2_ERR_LAS2, "svdLAS2: insufficient dimensions: " + std::to_string(dims));
        std::unique_ptr<Operator> op(create_operator(format, nrows, ncols, nnz, a, b, values, -1));
        solve(*op, dimensions, iterations, end_interval, kappa, seed, res);
        res->profile.t_total = wall_now() - t0;
    });
    if (rc != SVDLAS2_OK) {
        svdlas2_free(res);
        return rc;
    }
    *out = res;
    return rc;
}

} // namespace

extern "C" {

int svdlas2_abi_version(void) { return SVDLAS2_ABI_VERSION; }

int svdlas2_trim_caches(void) {
    return guarded([&] {
        host_pool_trim();
        device_pool_trim();
        plan_cache_trim();
    });
}

const char *svdlas2_last_error(void) { return g_last_error.c_str(); }

int svdlas2_csr(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *row_offsets, const uint64_t *col_indices,
                const double *values, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                double kappa, uint32_t random_seed, svdlas2_result **out) {
    return one_shot(SVDLAS2_FMT_CSR, nrows, ncols, nnz, row_offsets, col_indices, values, dimensions, iterations,
                    end_interval, kappa, random_seed, out);
}

int svdlas2_csc(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *col_offsets, const uint64_t *row_indices,
                const double *values, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                double kappa, uint32_t random_seed, svdlas2_result **out) {
    return one_shot(SVDLAS2_FMT_CSC, nrows, ncols, nnz, col_offsets, row_indices, values, dimensions, iterations,
                    end_interval, kappa, random_seed, out);
}

int svdlas2_coo(uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *row_indices, const uint64_t *col_indices,
                const double *values, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                double kappa, uint32_t random_seed, svdlas2_result **out) {
    return one_shot(SVDLAS2_FMT_COO, nrows, ncols, nnz, row_indices, col_indices, values, dimensions, iterations,
                    end_interval, kappa, random_seed, out);
}

void svdlas2_free(svdlas2_result *r) {
    if (!r) return;
    host_release(r->ut); host_release(r->vt); // pinned pool (falls back to free() for foreign pointers)
    std::free(r->s);
    std::free(r->alf); std::free(r->bet); std::free(r->ritz); std::free(r->bnd);
    std::free(r);
}

int svdlas2_operator_create(int format, uint64_t nrows, uint64_t ncols, uint64_t nnz, const uint64_t *a,
                            const uint64_t *b, const double *values, int device, svdlas2_operator **out) {
    if (!out) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "op is null");
    *out = nullptr;
    return guarded([&] { *out = reinterpret_cast<svdlas2_operator *>(create_operator(format, nrows, ncols, nnz, a, b, values, device)); });
}

void svdlas2_operator_destroy(svdlas2_operator *op_) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op) return;
    cudaSetDevice(op->device);
    if (op->stream) cudaStreamSynchronize(op->stream);
    delete op; // (the peer-memory ring belongs to the communicator)
}

int svdlas2_operator_solve(svdlas2_operator *op_, uint64_t dimensions, uint64_t iterations, const double end_interval[2],
                           double kappa, uint32_t random_seed, svdlas2_result **out) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || !out || !end_interval) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    svdlas2_result *res = nullptr;
    int rc = guarded([&] {
        res = new_result();
        solve(*op, dimensions, iterations, end_interval, kappa, random_seed, res);
    });
    if (rc != SVDLAS2_OK) {
        svdlas2_free(res);
        return rc;
    }
    *out = res;
    return rc;
}

int svdlas2_operator_shape(const svdlas2_operator *op_, uint64_t *m, uint64_t *n, uint64_t *nnz, int *transposed) {
    const Operator *op = reinterpret_cast<const Operator *>(op_);
    if (!op) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null operator");
    if (m) *m = op->M;
    if (n) *n = op->N;
    if (nnz) *nnz = op->B.nnz;
    if (transposed) *transposed = op->transposed ? 1 : 0;
    return SVDLAS2_OK;
}

int svdlas2_operator_opa(svdlas2_operator *op_, const double *x, double *y, int transposed) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || !x || !y) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        if (op->comm) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "opa is not available on a sharded operator");
        LAS2_CUDA(cudaSetDevice(op->device));
        // y = A x or A' x ; with B = A (not transposed) or B = A' (transposed): A = B or B'
        const bool use_b = (transposed != 0) == op->transposed; // A x -> B x when !op.transposed ; A'x -> B x when op.transposed
        const DeviceCsr &Mx = use_b ? op->B : op->BT;
        DevBuf<double> dx(Mx.cols ? Mx.cols : 1), dy(Mx.rows ? Mx.rows : 1);
        // the N side lives in the relabelled order on the device: permute on the way in (B x) or out (B' t)
        const std::vector<u32> &n2o = op->relabel.h_new2old;
        std::vector<double> hx;
        if (op->relabel.active && use_b) {
            hx.resize(Mx.cols);
            for (u64 i = 0; i < Mx.cols; i++) hx[i] = x[n2o[i]];
            x = hx.data();
        }
        LAS2_CUDA(cudaMemcpyAsync(dx.get(), x, Mx.cols * sizeof(double), cudaMemcpyHostToDevice, op->stream));
        LAS2_CUDA(cudaStreamSynchronize(op->stream));
        // poison the output (NaN pattern): a row the kernel forgets to write must not look plausible
        LAS2_CUDA(cudaMemsetAsync(dy.get(), 0xff, Mx.rows * sizeof(double), op->stream));
        spmv(Mx, dx.get(), dy.get(), op->stream);
        if (op->relabel.active && !use_b) {
            std::vector<double> hy(Mx.rows);
            LAS2_CUDA(cudaMemcpyAsync(hy.data(), dy.get(), Mx.rows * sizeof(double), cudaMemcpyDeviceToHost, op->stream));
            LAS2_CUDA(cudaStreamSynchronize(op->stream));
            for (u64 i = 0; i < Mx.rows; i++) y[n2o[i]] = hy[i];
            return;
        }
        LAS2_CUDA(cudaMemcpyAsync(y, dy.get(), Mx.rows * sizeof(double), cudaMemcpyDeviceToHost, op->stream));
        LAS2_CUDA(cudaStreamSynchronize(op->stream));
    });
}

int svdlas2_operator_opb(svdlas2_operator *op_, const double *x, double *y) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || !x || !y) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        LAS2_CUDA(cudaSetDevice(op->device));
        DevBuf<double> dx(op->N ? op->N : 1), dy(op->N ? op->N : 1), dt(op->M ? op->M : 1);
        const std::vector<u32> &n2o = op->relabel.h_new2old;
        std::vector<double> hx, hy;
        if (op->relabel.active) {
            hx.resize(op->N); hy.resize(op->N);
            for (u64 i = 0; i < op->N; i++) hx[i] = x[n2o[i]];
            x = hx.data();
        }
        LAS2_CUDA(cudaMemcpyAsync(dx.get(), x, op->N * sizeof(double), cudaMemcpyHostToDevice, op->stream));
        op->opb(dx.get(), dy.get(), dt.get(), nullptr);
        LAS2_CUDA(cudaMemcpyAsync(op->relabel.active ? hy.data() : y, dy.get(), op->N * sizeof(double), cudaMemcpyDeviceToHost, op->stream));
        LAS2_CUDA(cudaStreamSynchronize(op->stream));
        if (op->relabel.active)
            for (u64 i = 0; i < op->N; i++) y[n2o[i]] = hy[i];
    });
}

int svdlas2_operator_time_opb(svdlas2_operator *op_, int reps, int flush_l2, double *ms_per_opb, double *ms_spmv_b,
                              double *ms_spmv_bt) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || reps < 1) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        LAS2_CUDA(cudaSetDevice(op->device));
        cudaStream_t s = op->stream;
        DevBuf<double> dx(op->N ? op->N : 1), dy(op->N ? op->N : 1), dt(op->M ? op->M : 1);
        std::vector<double> hx(op->N);
        for (u64 i = 0; i < op->N; i++) hx[i] = 1.0 / (double)(1 + (i % 97));
        LAS2_CUDA(cudaMemcpyAsync(dx.get(), hx.data(), op->N * sizeof(double), cudaMemcpyHostToDevice, s));
        const size_t flush_bytes = (size_t)256 << 20; // > 126 MB L2
        if (flush_l2 && op->l2_flush.size() < flush_bytes) op->l2_flush.alloc(flush_bytes);
        cudaEvent_t e0, e1, e2;
        LAS2_CUDA(cudaEventCreate(&e0)); LAS2_CUDA(cudaEventCreate(&e1)); LAS2_CUDA(cudaEventCreate(&e2));
        for (int w = 0; w < 3; w++) op->opb(dx.get(), dy.get(), dt.get(), nullptr); // warm-up
        LAS2_CUDA(cudaStreamSynchronize(s));
        double tb = 0.0, tbt = 0.0;
        for (int i = 0; i < reps; i++) {
            if (flush_l2) LAS2_CUDA(cudaMemsetAsync(op->l2_flush.get(), i & 0xff, flush_bytes, s));
            LAS2_CUDA(cudaEventRecord(e0, s));
            spmv(op->B, dx.get(), dt.get(), s);
            LAS2_CUDA(cudaEventRecord(e1, s));
            if (op->ring) { // the sharded exchange as it runs inside a Lanczos step: peer-memory sum fused with the epilogue
                spmv(op->BT, dt.get(), op->ring->y_mine(), s);
                op->ring->reduce((op->N + 31) / 32 * 32, nullptr, 0.0, nullptr, nullptr, s);
            } else {
                spmv(op->BT, dt.get(), dy.get(), s);
                if (op->comm) op->comm->allreduce_sum(dy.get(), op->N, s);
            }
            LAS2_CUDA(cudaEventRecord(e2, s));
            LAS2_CUDA(cudaStreamSynchronize(s));
            float a = 0.f, b = 0.f;
            LAS2_CUDA(cudaEventElapsedTime(&a, e0, e1));
            LAS2_CUDA(cudaEventElapsedTime(&b, e1, e2));
            tb += a; tbt += b;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (ms_spmv_b) *ms_spmv_b = tb / reps;
        if (ms_spmv_bt) *ms_spmv_bt = tbt / reps;
        if (ms_per_opb) *ms_per_opb = (tb + tbt) / reps;
    });
}

int svdlas2_test_time_panel(uint64_t n, uint32_t ncols, int reps, int device, double *ms_dots, double *ms_update) {
    if (n == 0 || ncols == 0 || reps < 1) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        if (device >= 0) LAS2_CUDA(cudaSetDevice(device));
        const u64 ld = (n + 31) / 32 * 32;
        DevBuf<double> Q((size_t)ld * ncols), a(ld), b(ld), parts(3 * (size_t)kMaxParts);
        cudaStream_t s;
        LAS2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        // small entries so that repeated passes stay finite: the values do not matter for the timing
        LAS2_CUDA(cudaMemsetAsync(Q.get(), 0, (size_t)ld * ncols * sizeof(double), s));
        LAS2_CUDA(cudaMemsetAsync(a.get(), 0, ld * sizeof(double), s));
        LAS2_CUDA(cudaMemsetAsync(b.get(), 0, ld * sizeof(double), s));
        PanelWork pw;
        PartialSlot sa{parts.get(), 0}, sb{parts.get() + kMaxParts, 0}, sd{parts.get() + 2 * kMaxParts, 0};
        cudaEvent_t e0, e1, e2;
        LAS2_CUDA(cudaEventCreate(&e0)); LAS2_CUDA(cudaEventCreate(&e1)); LAS2_CUDA(cudaEventCreate(&e2));
        double td = 0.0, tu = 0.0;
        for (int i = -2; i < reps; i++) {
            panel_cgs2(Q.get(), ld, n, 0, ncols, a.get(), b.get(), pw, &sa, &sb, &sd, s, nullptr, e0, e1, e2);
            LAS2_CUDA(cudaStreamSynchronize(s));
            float x = 0.f, y = 0.f;
            LAS2_CUDA(cudaEventElapsedTime(&x, e0, e1));
            LAS2_CUDA(cudaEventElapsedTime(&y, e1, e2));
            if (i >= 0) { td += x; tu += y; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        cudaStreamDestroy(s);
        if (ms_dots) *ms_dots = td / reps;
        if (ms_update) *ms_update = tu / reps;
    });
}

int svdlas2_test_time_ritz(uint64_t n, uint32_t js, uint32_t d, int reps, int device, double ms[2], double *max_diff, double *max_err) {
    if (n == 0 || js == 0 || d == 0 || reps < 1 || !ms) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        if (device >= 0) LAS2_CUDA(cudaSetDevice(device));
        const u64 ld = (n + 31) / 32 * 32;
        cudaStream_t s;
        LAS2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        struct SDel { cudaStream_t s; ~SDel() { cudaStreamDestroy(s); } } sdel{s};
        // pseudo-random entries in [-1, 1): cheap on the host for C, hashed on the host for a strip of Q and replicated
        std::vector<double> hC((size_t)js * d), hQ((size_t)ld * js, 0.0);
        u64 st = 0x243F6A8885A308D3ull;
        auto rnd = [&] { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) * (2.0 / 9007199254740992.0) - 1.0; };
        for (auto &v : hC) v = rnd();
        for (u32 k = 0; k < js; k++)
            for (u64 i = 0; i < n; i++) hQ[(size_t)k * ld + i] = rnd();
        DevBuf<double> Q((size_t)ld * js), C((size_t)js * d), V0((size_t)ld * d), V1((size_t)ld * d);
        LAS2_CUDA(cudaMemcpyAsync(Q.get(), hQ.data(), hQ.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        LAS2_CUDA(cudaMemcpyAsync(C.get(), hC.data(), hC.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        cudaEvent_t e0, e1;
        LAS2_CUDA(cudaEventCreate(&e0)); LAS2_CUDA(cudaEventCreate(&e1));
        for (int variant = 0; variant < 2; variant++) {
            g_ritz_variant = variant;
            double *V = variant ? V1.get() : V0.get();
            LAS2_CUDA(cudaMemsetAsync(V, 0xff, (size_t)ld * d * sizeof(double), s));
            ritz_contract(Q.get(), ld, n, js, C.get(), d, V, s, nullptr); // warm-up
            LAS2_CUDA(cudaEventRecord(e0, s));
            for (int r = 0; r < reps; r++) ritz_contract(Q.get(), ld, n, js, C.get(), d, V, s, nullptr);
            LAS2_CUDA(cudaEventRecord(e1, s));
            LAS2_CUDA(cudaStreamSynchronize(s));
            float t = 0.f;
            LAS2_CUDA(cudaEventElapsedTime(&t, e0, e1));
            ms[variant] = t / reps;
        }
        g_ritz_variant = -1;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        std::vector<double> h0((size_t)ld * d), h1((size_t)ld * d);
        LAS2_CUDA(cudaMemcpy(h0.data(), V0.get(), h0.size() * sizeof(double), cudaMemcpyDeviceToHost));
        LAS2_CUDA(cudaMemcpy(h1.data(), V1.get(), h1.size() * sizeof(double), cudaMemcpyDeviceToHost));
        double vmax = 0.0, dmax = 0.0, emax = 0.0;
        for (u32 c = 0; c < d; c++)
            for (u64 i = 0; i < n; i++) {
                const double a = h0[(size_t)c * ld + i], b = h1[(size_t)c * ld + i];
                vmax = std::max(vmax, std::fabs(a));
                dmax = std::max(dmax, std::fabs(a - b));
            }
        // reference in long double on a sample of entries
        const u64 stride_i = std::max<u64>(1, n / 97);
        const u32 stride_c = std::max<u32>(1, d / 13);
        for (u32 c = 0; c < d; c += stride_c)
            for (u64 i = 0; i < n; i += stride_i) {
                long double acc = 0.0L;
                for (u32 k = 0; k < js; k++) acc += (long double)hQ[(size_t)k * ld + i] * (long double)hC[(size_t)k * d + c];
                emax = std::max(emax, (double)fabsl(acc - (long double)h1[(size_t)c * ld + i]));
            }
        if (max_diff) *max_diff = vmax > 0 ? dmax / vmax : dmax;
        if (max_err) *max_err = vmax > 0 ? emax / vmax : emax;
    });
}

int svdlas2_test_ql_replay(uint64_t n, double *d, double *e, double *z, int variant, int reps, int device, double *ms,
                           uint64_t *n_rotations) {
    if (n == 0 || !d || !e || !z || reps < 1) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        if (device >= 0) LAS2_CUDA(cudaSetDevice(device));
        QlPlan plan;
        const int rc = ql_rotations(n, d, e, plan);
        if (rc) throw Error(rc, "imtql2 did not converge");
        if (n_rotations) *n_rotations = plan.rotations();
        cudaStream_t s;
        LAS2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        struct SDel { cudaStream_t s; ~SDel() { cudaStreamDestroy(s); } } sdel{s};
        const u32 js = (u32)n, ldz = (u32)((n + 31) / 32 * 32);
        DevBuf<double> zT((size_t)ldz * js);
        struct VReset { ~VReset() { g_ql_replay_variant = -1; } } vreset;
        g_ql_replay_variant = variant;
        LAS2_CUDA(cudaMemsetAsync(zT.get(), 0xff, (size_t)ldz * js * sizeof(double), s)); // NaN: every entry must be written
        ql_replay(plan, js, ldz, zT.get(), s, nullptr, nullptr); // (synchronises the stream itself)
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 1; r < reps; r++) ql_replay(plan, js, ldz, zT.get(), s, nullptr, nullptr);
        LAS2_CUDA(cudaStreamSynchronize(s));
        if (ms) *ms = reps > 1 ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (reps - 1) : 0.0;
        std::vector<double> h((size_t)ldz * js);
        LAS2_CUDA(cudaMemcpy(h.data(), zT.get(), h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (u64 row = 0; row < n; row++)
            for (u64 k = 0; k < n; k++) z[row * n + k] = h[(size_t)plan.column_of[k] * ldz + row];
    });
}

int svdlas2_test_device_parking(int device, double out[4]) {
    if (!out) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    return guarded([&] {
        if (device >= 0) LAS2_CUDA(cudaSetDevice(device));
        device_pool_trim();
        size_t free_b = 0, total_b = 0;
        LAS2_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t small = free_b / 10 * 3, large = free_b / 10 * 8; // together more than the device has
        double st[2];
        void *a = device_alloc(small);
        LAS2_CUDA(cudaMemset(a, 0, 1 << 20));
        device_free(a);
        device_parked_stats(st);
        out[0] = st[1];
        void *b = device_alloc(small - small / 16); // fits the parked block with less than 1/8 of slack
        out[1] = b == a ? 1.0 : 0.0;
        device_free(b);
        void *c = device_alloc(large); // only possible once the parked block has gone back to the driver
        LAS2_CUDA(cudaMemset(c, 0, 1 << 20));
        out[2] = c ? 1.0 : 0.0;
        device_free(c);
        device_pool_trim();
        device_parked_stats(st);
        out[3] = st[1];
    });
}

int svdlas2_test_operator_layout(const svdlas2_operator *op_, double out[8]) {
    const Operator *op = reinterpret_cast<const Operator *>(op_);
    if (!op || !out) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    auto mode = [](const ChunkStream &K) -> double {
        if (!K.active) return -1.0;
        if (K.slabs) return K.slab_local ? 2.0 : 3.0;
        return K.hot_entries > 0 ? 1.0 : 0.0;
    };
    out[0] = op->relabel.active ? 1.0 : 0.0;
    out[1] = op->B.hot_fraction;
    out[2] = (double)op->B.hot_entries;
    out[3] = mode(op->B.chk);
    out[4] = mode(op->BT.chk);
    out[5] = mode(op->BT.chk2);
    out[6] = (double)op->BT.chk2.row0;
    out[7] = (double)op->BT.chk.nslabs;
    return SVDLAS2_OK;
}

int svdlas2_test_spmm4(svdlas2_operator *op_, uint32_t nv, const double *X, double *Y, double *sumsq) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || !X || !Y || nv < 1 || nv > 4) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument");
    return guarded([&] {
        LAS2_CUDA(cudaSetDevice(op->device));
        if (!spmm4_supported(op->B)) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "B is not on the (plain) chunk stream");
        cudaStream_t s = op->stream;
        const u64 N = op->N, M = op->M, ld = (N + 31) / 32 * 32;
        DevBuf<double> dx((size_t)nv * ld), dy((size_t)nv * M), ds(nv);
        LAS2_CUDA(cudaMemsetAsync(dx.get(), 0, (size_t)nv * ld * sizeof(double), s));
        std::vector<double> hx;
        if (op->relabel.active) {
            hx.resize((size_t)nv * N);
            for (uint32_t r = 0; r < nv; r++)
                for (u64 i = 0; i < N; i++) hx[(size_t)r * N + i] = X[(size_t)r * N + op->relabel.h_new2old[i]];
            X = hx.data();
        }
        LAS2_CUDA(cudaMemcpy2DAsync(dx.get(), ld * sizeof(double), X, N * sizeof(double), N * sizeof(double), nv, cudaMemcpyHostToDevice, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
        LAS2_CUDA(cudaMemsetAsync(dy.get(), 0xff, (size_t)nv * M * sizeof(double), s));
        Spmm4Work w4;
        const int np = spmm4(op->B, dx.get(), ld, nv, dy.get(), M, w4, s);
        GatherList g;
        g.count = (int)nv;
        for (uint32_t r = 0; r < nv; r++) { g.parts[r] = w4.parts.get() + (size_t)r * kMaxParts; g.nparts[r] = np; }
        gather_scalars(g, ds.get(), s);
        LAS2_CUDA(cudaMemcpyAsync(Y, dy.get(), (size_t)nv * M * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (sumsq) LAS2_CUDA(cudaMemcpyAsync(sumsq, ds.get(), nv * sizeof(double), cudaMemcpyDeviceToHost, s));
        LAS2_CUDA(cudaStreamSynchronize(s));
    });
}

int svdlas2_test_ring_emulation(int G, uint64_t n, const double *ys, const double *qprev, double rnm, const double *q,
                                int two_shot, int ctas_per_rank, double *out, double *dots) {
    if (!ys || !out || n == 0 || n % 32) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "bad argument (n must be a multiple of 32)");
    return guarded([&] {
        cudaStream_t s;
        LAS2_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        struct SDel { cudaStream_t s; ~SDel() { cudaStreamDestroy(s); } } sdel{s};
        DevBuf<double> dys((size_t)G * n), dout((size_t)G * n), dq(n), dqp(n), dparts((size_t)G * G * ctas_per_rank + 1), dd(G);
        LAS2_CUDA(cudaMemcpyAsync(dys.get(), ys, (size_t)G * n * sizeof(double), cudaMemcpyHostToDevice, s));
        if (q) LAS2_CUDA(cudaMemcpyAsync(dq.get(), q, n * sizeof(double), cudaMemcpyHostToDevice, s));
        if (qprev) LAS2_CUDA(cudaMemcpyAsync(dqp.get(), qprev, n * sizeof(double), cudaMemcpyHostToDevice, s));
        LAS2_CUDA(cudaMemsetAsync(dout.get(), 0xff, (size_t)G * n * sizeof(double), s));
        int np = 0;
        ring_emulate(G, n, dys.get(), qprev ? dqp.get() : nullptr, rnm, q ? dq.get() : nullptr, two_shot, ctas_per_rank, dout.get(),
                     dparts.get(), &np, s);
        LAS2_CUDA(cudaMemcpyAsync(out, dout.get(), (size_t)G * n * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (q && dots) {
            for (int g = 0; g < G; g += 8) { // the canonical sum every consumer kernel forms from the partials
                GatherList gl;
                gl.count = std::min(8, G - g);
                for (int k = 0; k < gl.count; k++) { gl.parts[k] = dparts.get() + (size_t)(g + k) * G * ctas_per_rank; gl.nparts[k] = np; }
                gather_scalars(gl, dd.get() + g, s);
            }
            LAS2_CUDA(cudaMemcpyAsync(dots, dd.get(), G * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        LAS2_CUDA(cudaStreamSynchronize(s));
    });
}

int svdlas2_operator_info(const svdlas2_operator *op_, const char *key, double *value) {
    const Operator *op = reinterpret_cast<const Operator *>(op_);
    if (!op || !key || !value) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    const PeerRing *R = op->ring;
    const u64 ld = (op->N + 31) / 32 * 32;
    if (!std::strcmp(key, "relabelled")) { *value = op->relabel.active ? 1.0 : 0.0; return SVDLAS2_OK; }
    if (!std::strcmp(key, "ring")) { *value = R ? 1.0 : 0.0; return SVDLAS2_OK; }
    if (!std::strcmp(key, "ring_calls")) { *value = R ? (double)R->calls : 0.0; return SVDLAS2_OK; }
    if (!std::strcmp(key, "ring_two_shot_calls")) { *value = R ? (double)R->two_shot_calls : 0.0; return SVDLAS2_OK; }
    if (!std::strcmp(key, "ring_bytes_per_call")) {
        // one-shot: reads G-1 whole partial vectors; two-shot: reads G-1 slices of N/G and pushes its slice to G-1 peers
        double b = 0.0;
        if (R) b = (ld * 8 > R->two_shot_min_bytes) ? 2.0 * (R->G - 1) * (double)ld * 8.0 / R->G : (double)(R->G - 1) * (double)ld * 8.0;
        *value = b;
        return SVDLAS2_OK;
    }
    if (!std::strcmp(key, "hot_fraction")) { *value = op->B.hot_fraction; return SVDLAS2_OK; }
    return fail(SVDLAS2_ERR_INVALID_ARGUMENT, std::string("unknown key ") + key);
}

int svdlas2_operator_set(svdlas2_operator *op_, const char *knob, int64_t value) {
    Operator *op = reinterpret_cast<Operator *>(op_);
    if (!op || !knob) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null argument");
    if (!std::strcmp(knob, "profile")) { op->profile = value != 0; return SVDLAS2_OK; }
    if (!std::strcmp(knob, "nside_on_rank0_only")) { op->nside_mode = value != 0 ? 1 : 0; return SVDLAS2_OK; }
    if (!std::strcmp(knob, "nside_mode")) {
        if (value < 0 || value > 2) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "nside_mode must be 0, 1 or 2");
        op->nside_mode = (int)value;
        return SVDLAS2_OK;
    }
    // SpMV tuning: stored in THIS operator's two matrices (no process-wide state)
    auto both = [&](int SpmvTuning::*field) { op->B.tune.*field = (int)value; op->BT.tune.*field = (int)value; return SVDLAS2_OK; };
    if (!std::strcmp(knob, "spmv_variant")) return both(&SpmvTuning::variant);
    if (!std::strcmp(knob, "spmv_hot")) return both(&SpmvTuning::hot_entries);
    if (!std::strcmp(knob, "spmv_smem_pad_kb")) return both(&SpmvTuning::smem_pad_kb);
    if (!std::strcmp(knob, "spmv_pdl")) return both(&SpmvTuning::pdl);
    if (!std::strcmp(knob, "spmv_shape")) return both(&SpmvTuning::shape);
    if (!std::strcmp(knob, "spmv_carveout")) return both(&SpmvTuning::carveout_pct);
    return fail(SVDLAS2_ERR_INVALID_ARGUMENT, std::string("unknown knob ") + knob);
}

int svdlas2_comm_unique_id(uint8_t id[SVDLAS2_COMM_ID_BYTES]) {
    if (!id) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "null id");
    return guarded([&] { Comm::unique_id(id); });
}

int svdlas2_comm_create(int rank, int world_size, const uint8_t id[SVDLAS2_COMM_ID_BYTES], int device,
                        svdlas2_comm **comm) {
    if (!comm) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "comm is null");
    *comm = nullptr;
    return guarded([&] {
        if (world_size > 1 && !id) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "null communicator id");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            throw Error(SVDLAS2_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
        }
        if (device >= 0) LAS2_CUDA(cudaSetDevice(device));
        *comm = reinterpret_cast<svdlas2_comm *>(Comm::create(rank, world_size, id));
    });
}

void svdlas2_comm_destroy(svdlas2_comm *comm) { delete reinterpret_cast<Comm *>(comm); }

int svdlas2_operator_create_shard(uint64_t m_global, uint64_t n, uint64_t row_begin, uint64_t row_end,
                                  uint64_t local_nnz, const uint64_t *local_offsets, const uint64_t *local_indices,
                                  const double *values, int b_is_a_transposed, uint64_t global_nnz, svdlas2_comm *comm,
                                  int device, svdlas2_operator **out) {
    if (!out) return fail(SVDLAS2_ERR_INVALID_ARGUMENT, "op is null");
    *out = nullptr;
    return guarded([&] {
        if (row_end < row_begin || row_end > m_global) throw Error(SVDLAS2_ERR_INVALID_ARGUMENT, "bad row range");
        std::unique_ptr<Operator> op(new Operator());
        select_device(device, *op);
        const double t0 = wall_now();
        op->transposed = b_is_a_transposed != 0;
        op->a_rows = op->transposed ? n : m_global;
        op->a_cols = op->transposed ? m_global : n;
        op->nnz_global = global_nnz;
        op->M_global = m_global;
        op->M = row_end - row_begin;
        op->row_begin = row_begin;
        op->N = n;
        Comm *c = reinterpret_cast<Comm *>(comm);
        op->comm = (c && c->world() > 1) ? c : nullptr;
        build_operator_from_csr(op->M, n, local_nnz, local_offsets, local_indices, values, op->stream, op->B, op->BT,
                                op->upload, op->relabel, op->comm);
        // collective; every rank reaches the same verdict.  Without peer memory the exchange stays on ncclAllReduce.
        if (op->comm) op->ring = op->comm->vector_ring((n + 31) / 32 * 32);
        if (op->ring) {
            op->ring->reset_vectors(op->stream);
            LAS2_CUDA(cudaStreamSynchronize(op->stream));
        }
        op->t_upload = wall_now() - t0;
        *out = reinterpret_cast<svdlas2_operator *>(op.release());
    });
}

} // extern "C"
