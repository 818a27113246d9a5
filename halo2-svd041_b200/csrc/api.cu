// C ABI of libh2svd_b200 (include/h2svd_b200.h): handle management, argument checking, host<->device
// staging for the host-pointer entry points.  No CPU fallback anywhere: every entry point either runs
// the CUDA kernels or fails with a negative code.
#include <stdarg.h>
#include <string.h>

#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace h2svd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? H2SVD_ENOMEM : H2SVD_ECUDA;
}

int ws_reserve(h2svd_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return H2SVD_OK;
    if (ctx->capturing) {
        set_error("graph capture: the workspace would have to grow -- run the same calls once before h2svd_graph_begin");
        return H2SVD_EINVAL;
    }
    // the old block may still be in use by queued work: drain first
    H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->side_stream));
    if (ctx->ws) H2SVD_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    H2SVD_CUDA(cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
    ctx->ws_generation++;
    return H2SVD_OK;
}

// grow-only side buffers (operand planes, stream-K partials): same rules as ws_reserve
int ws_grow(h2svd_ctx* ctx, void** buf, size_t* cur, size_t need) {
    if (*cur >= need) return H2SVD_OK;
    if (ctx->capturing) {
        set_error("graph capture: a workspace would have to grow -- run the same calls once before h2svd_graph_begin");
        return H2SVD_EINVAL;
    }
    H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->side_stream));
    if (*buf) H2SVD_CUDA(cudaFree(*buf));
    *buf = nullptr;
    *cur = 0;
    H2SVD_CUDA(cudaMalloc(buf, need));
    *cur = need;
    ctx->ws_generation++;
    return H2SVD_OK;
}

int rescale_params(int P, int lb, int S, int A, int* n_d, int* n_r);  // rescale.cu

// bump allocator over the workspace (all offsets 256-byte aligned)
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    template <typename T>
    T* take(size_t count) {
        T* p = (T*)(base + off);
        off += (count * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
    static size_t need(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};

static int check_flag(h2svd_ctx* ctx, const char* what) {
    int flag = 0;
    H2SVD_CUDA(cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        H2SVD_CUDA(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
        set_error("%s: operand out of range / non-canonical field element", what);
        return H2SVD_ERANGE;
    }
    return H2SVD_OK;
}

#define REQUIRE(cond, msg)              \
    do {                                \
        if (!(cond)) {                  \
            h2svd::set_error("%s", msg); \
            return H2SVD_EINVAL;        \
        }                               \
    } while (0)

static int h2d(h2svd_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return H2SVD_OK;
    H2SVD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return H2SVD_OK;
}
static int d2h(h2svd_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return H2SVD_OK;
    H2SVD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return H2SVD_OK;
}

// Freivalds on device pointers; scratch = k + 2n elements for (b v), (c_s v), (a b v) row totals
static int freivalds_dev(h2svd_ctx* ctx, const Fr* a, const Fr* b, const Fr* cs, const Fr* gamma, size_t n,
                         size_t k, size_t m, Fr* powers, Fr* pcv, Fr* pbv, Fr* pabv, Fr* diff, Fr* is_zero,
                         Fr* inv, Fr* scratch) {
    Fr* bv = scratch;
    Fr* csv = scratch + k;
    Fr* abv = scratch + k + n;
    H2SVD_TRY(launch_gamma_powers(ctx, gamma, m, powers));                                  // reference :316-326
    H2SVD_TRY(launch_mat_vec_prefix2(ctx, cs, n, pcv, csv, b, k, pbv, bv, powers, m));      // :335, :336 (one launch)
    H2SVD_TRY(launch_mat_vec_prefix(ctx, a, bv, n, k, 0, pabv, abv));                       // :337
    H2SVD_TRY(launch_is_equal(ctx, csv, abv, n, diff, is_zero, inv));                       // :339-341
    return H2SVD_OK;
}

}  // namespace h2svd

namespace h2svd {
// after a failure inside a pipelined host call: nothing may still be writing the caller's buffers, no stale flag
static void drain_after_error(h2svd_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->side_stream);
    cudaMemset(ctx->d_flag, 0, sizeof(int));
    cudaGetLastError();
}
}  // namespace h2svd
using namespace h2svd;

extern "C" {

const char* h2svd_last_error(void) { return g_err; }
const char* h2svd_version(void) { return "h2svd_b200 0.1 (sm_100a)"; }

int h2svd_create(h2svd_ctx** out, int device, void* stream) {
    REQUIRE(out != nullptr, "h2svd_create: out is null");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("h2svd_create: no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return H2SVD_ENODEV;
    }
    if (device < 0) H2SVD_CUDA(cudaGetDevice(&device));
    REQUIRE(device < ndev, "h2svd_create: device index out of range");
    H2SVD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    H2SVD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("h2svd_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return H2SVD_ENODEV;
    }
    h2svd_ctx* ctx = new (std::nothrow) h2svd_ctx();
    if (!ctx) return H2SVD_ENOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            delete ctx;
            return cuda_fail(se, "cudaStreamCreate", __FILE__, __LINE__);
        }
        ctx->owns_stream = true;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&ctx->d_flag, 4 * sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_flag, 0, 4 * sizeof(int)) != cudaSuccess) {
        h2svd_destroy(ctx);
        return cuda_fail(cudaGetLastError(), "handle setup", __FILE__, __LINE__);
    }
    ctx->d_mode = ctx->d_flag + 1;
    for (int i = 0; i < 8; i++) {
        if (cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            h2svd_destroy(ctx);
            return cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
        }
    }
    *out = ctx;
    return H2SVD_OK;
}

void h2svd_destroy(h2svd_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->side_stream) {
        cudaStreamSynchronize(ctx->side_stream);
        cudaStreamDestroy(ctx->side_stream);
    }
    for (int i = 0; i < 8; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->sk_ws) cudaFree(ctx->sk_ws);
    if (ctx->kara_ws) cudaFree(ctx->kara_ws);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    if (ctx->d_timeline) cudaFree(ctx->d_timeline);
    if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int h2svd_sync(h2svd_ctx* ctx) {
    REQUIRE(ctx, "h2svd_sync: null handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return check_flag(ctx, "h2svd_sync");
}
void* h2svd_stream(h2svd_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int h2svd_device(h2svd_ctx* ctx) { return ctx ? ctx->device : -1; }
int h2svd_sm_count(h2svd_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t h2svd_launch_count(h2svd_ctx* ctx) { return ctx ? ctx->launches : 0; }

/* ---- K1 ---- */
int h2svd_fr_matmul_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* c, size_t n, size_t k,
                        size_t m, int b_transposed) {
    REQUIRE(ctx && a && b && c, "fr_matmul: null argument");
    REQUIRE(n >= 1 && k >= 1 && m >= 1, "fr_matmul: empty matrix");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const Fr* bp = as_fr(b);
    if (b_transposed) {
        H2SVD_TRY(ws_reserve(ctx, k * m * sizeof(Fr)));
        H2SVD_TRY(launch_transpose(ctx, as_fr(b), (Fr*)ctx->ws, m, k));  // b is m x k -> k x m
        bp = (const Fr*)ctx->ws;
    }
    return launch_fr_matmul(ctx, as_fr(a), bp, as_fr(c), n, k, m);
}

int h2svd_fr_matmul(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* c, size_t n, size_t k,
                    size_t m, int b_transposed) {
    REQUIRE(ctx && a && b && c, "fr_matmul: null argument");
    REQUIRE(n >= 1 && k >= 1 && m >= 1, "fr_matmul: empty matrix");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t need = Carver::need(n * k * sizeof(Fr)) + 2 * Carver::need(k * m * sizeof(Fr)) +
                        Carver::need(n * m * sizeof(Fr));
    H2SVD_TRY(ws_reserve(ctx, need));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(n * k);
    Fr* db = cv.take<Fr>(k * m);
    Fr* dbt = cv.take<Fr>(k * m);
    Fr* dc = cv.take<Fr>(n * m);
    H2SVD_TRY(h2d(ctx, da, a, n * k * sizeof(Fr)));
    H2SVD_TRY(h2d(ctx, db, b, k * m * sizeof(Fr)));
    H2SVD_TRY(launch_check_canonical(ctx, da, n * k, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, db, k * m, ctx->d_flag));
    const Fr* bp = db;
    if (b_transposed) {
        H2SVD_TRY(launch_transpose(ctx, db, dbt, m, k));
        bp = dbt;
    }
    H2SVD_TRY(launch_fr_matmul(ctx, da, bp, dc, n, k, m));
    H2SVD_TRY(d2h(ctx, c, dc, n * m * sizeof(Fr)));
    return check_flag(ctx, "fr_matmul");
}

/* ---- K2/K3 ---- */
int h2svd_gamma_powers_dev(h2svd_ctx* ctx, const h2svd_fr* gamma, size_t d, h2svd_fr* out) {
    REQUIRE(ctx && gamma && out, "gamma_powers: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_gamma_powers(ctx, as_fr(gamma), d, as_fr(out));
}
int h2svd_mat_vec_prefix_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t len,
                             h2svd_fr* out_prefix) {
    REQUIRE(ctx && a && v && out_prefix, "mat_vec_prefix: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_mat_vec_prefix(ctx, as_fr(a), as_fr(v), rows, len, 0, as_fr(out_prefix), nullptr);
}
int h2svd_mat_vec_prefix_totals_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t len,
                                    h2svd_fr* out_prefix, h2svd_fr* out_totals) {
    REQUIRE(ctx && a && v && out_prefix, "mat_vec_prefix_totals: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_mat_vec_prefix(ctx, as_fr(a), as_fr(v), rows, len, 0, as_fr(out_prefix),
                                 out_totals ? as_fr(out_totals) : nullptr);
}
int h2svd_mat_vec_prefix_pair_dev(h2svd_ctx* ctx, const h2svd_fr* a0, size_t rows0, h2svd_fr* out_prefix0,
                                  h2svd_fr* out_totals0, const h2svd_fr* a1, size_t rows1, h2svd_fr* out_prefix1,
                                  h2svd_fr* out_totals1, const h2svd_fr* v, size_t len) {
    REQUIRE(ctx && v, "mat_vec_prefix_pair: null argument");
    REQUIRE(rows0 == 0 || (a0 && out_prefix0), "mat_vec_prefix_pair: null argument (first matrix)");
    REQUIRE(rows1 == 0 || (a1 && out_prefix1), "mat_vec_prefix_pair: null argument (second matrix)");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_mat_vec_prefix2(ctx, as_fr(a0), rows0, as_fr(out_prefix0), out_totals0 ? as_fr(out_totals0) : nullptr,
                                  as_fr(a1), rows1, as_fr(out_prefix1), out_totals1 ? as_fr(out_totals1) : nullptr,
                                  as_fr(v), len);
}
int h2svd_gather_dev(h2svd_ctx* ctx, const h2svd_fr* src, size_t count, size_t stride, size_t offset,
                     h2svd_fr* out) {
    REQUIRE(ctx && src && out, "gather: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_gather(ctx, as_fr(src), count, stride, offset, as_fr(out));
}
int h2svd_is_equal_witness_dev(h2svd_ctx* ctx, const h2svd_fr* x, const h2svd_fr* y, size_t count,
                               h2svd_fr* diff, h2svd_fr* is_zero, h2svd_fr* inv) {
    REQUIRE(ctx && x && y && diff && is_zero && inv, "is_equal_witness: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_is_equal(ctx, as_fr(x), as_fr(y), count, as_fr(diff), as_fr(is_zero), as_fr(inv));
}

int h2svd_freivalds_witness_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* c_s,
                                const h2svd_fr* gamma, size_t n, size_t k, size_t m, h2svd_fr* powers,
                                h2svd_fr* prefix_cv, h2svd_fr* prefix_bv, h2svd_fr* prefix_abv, h2svd_fr* diff,
                                h2svd_fr* is_zero, h2svd_fr* inv) {
    REQUIRE(ctx && a && b && c_s && gamma && powers && prefix_cv && prefix_bv && prefix_abv && diff && is_zero &&
                inv,
            "freivalds_witness: null argument");
    REQUIRE(n >= 1 && k >= 1 && m >= 1, "freivalds_witness: empty matrix");  // reference :307-310
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    H2SVD_TRY(ws_reserve(ctx, (k + 2 * n) * sizeof(Fr)));
    return freivalds_dev(ctx, as_fr(a), as_fr(b), as_fr(c_s), as_fr(gamma), n, k, m, as_fr(powers),
                         as_fr(prefix_cv), as_fr(prefix_bv), as_fr(prefix_abv), as_fr(diff), as_fr(is_zero),
                         as_fr(inv), (Fr*)ctx->ws);
}

int h2svd_freivalds_witness(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* c_s,
                            const h2svd_fr* gamma, size_t n, size_t k, size_t m, h2svd_fr* powers,
                            h2svd_fr* prefix_cv, h2svd_fr* prefix_bv, h2svd_fr* prefix_abv, h2svd_fr* diff,
                            h2svd_fr* is_zero, h2svd_fr* inv) {
    REQUIRE(ctx && a && b && c_s && gamma && powers && prefix_cv && prefix_bv && prefix_abv && diff && is_zero &&
                inv,
            "freivalds_witness: null argument");
    REQUIRE(n >= 1 && k >= 1 && m >= 1, "freivalds_witness: empty matrix");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t F = sizeof(Fr);
    const size_t need = Carver::need(n * k * F) + Carver::need(k * m * F) + 2 * Carver::need(n * m * F) +
                        Carver::need(F) + Carver::need(m * F) + Carver::need(k * m * F) +
                        Carver::need(n * k * F) + 3 * Carver::need(n * F) + Carver::need((k + 2 * n) * F);
    H2SVD_TRY(ws_reserve(ctx, need));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(n * k);
    Fr* db = cv.take<Fr>(k * m);
    Fr* dcs = cv.take<Fr>(n * m);
    Fr* dg = cv.take<Fr>(1);
    Fr* dpow = cv.take<Fr>(m);
    Fr* dpcv = cv.take<Fr>(n * m);
    Fr* dpbv = cv.take<Fr>(k * m);
    Fr* dpabv = cv.take<Fr>(n * k);
    Fr* ddiff = cv.take<Fr>(n);
    Fr* dz = cv.take<Fr>(n);
    Fr* dinv = cv.take<Fr>(n);
    Fr* scratch = cv.take<Fr>(k + 2 * n);
    H2SVD_TRY(h2d(ctx, da, a, n * k * F));
    H2SVD_TRY(h2d(ctx, db, b, k * m * F));
    H2SVD_TRY(h2d(ctx, dcs, c_s, n * m * F));
    H2SVD_TRY(h2d(ctx, dg, gamma, F));
    H2SVD_TRY(launch_check_canonical(ctx, da, n * k, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, db, k * m, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dcs, n * m, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dg, 1, ctx->d_flag));
    H2SVD_TRY(freivalds_dev(ctx, da, db, dcs, dg, n, k, m, dpow, dpcv, dpbv, dpabv, ddiff, dz, dinv, scratch));
    H2SVD_TRY(d2h(ctx, powers, dpow, m * F));
    H2SVD_TRY(d2h(ctx, prefix_cv, dpcv, n * m * F));
    H2SVD_TRY(d2h(ctx, prefix_bv, dpbv, k * m * F));
    H2SVD_TRY(d2h(ctx, prefix_abv, dpabv, n * k * F));
    H2SVD_TRY(d2h(ctx, diff, ddiff, n * F));
    H2SVD_TRY(d2h(ctx, is_zero, dz, n * F));
    H2SVD_TRY(d2h(ctx, inv, dinv, n * F));
    return check_flag(ctx, "freivalds_witness");
}

int h2svd_mat_vec_prefix(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t len,
                         h2svd_fr* out_prefix) {
    REQUIRE(ctx && a && v && out_prefix, "mat_vec_prefix: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t cnt = rows * len, F = sizeof(Fr);
    if (cnt == 0) return H2SVD_OK;
    H2SVD_TRY(ws_reserve(ctx, 2 * Carver::need(cnt * F) + Carver::need(len * F)));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(cnt);
    Fr* dv = cv.take<Fr>(len);
    Fr* dout = cv.take<Fr>(cnt);
    H2SVD_TRY(h2d(ctx, da, a, cnt * F));
    H2SVD_TRY(h2d(ctx, dv, v, len * F));
    H2SVD_TRY(launch_check_canonical(ctx, da, cnt, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dv, len, ctx->d_flag));
    H2SVD_TRY(launch_mat_vec_prefix(ctx, da, dv, rows, len, 0, dout, nullptr));
    H2SVD_TRY(d2h(ctx, out_prefix, dout, cnt * F));
    return check_flag(ctx, "mat_vec_prefix");
}

/* ---- K4 ---- */
int h2svd_rescale_witness_count(int precision_bits, int lookup_bits, int shift_bits, int a_num_bits) {
    const int w = rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr);
    if (w < 0) {
        set_error("rescale_witness_count: parameters out of range");
        return H2SVD_EINVAL;
    }
    return w;
}

int h2svd_rescale_witness_dev(h2svd_ctx* ctx, const h2svd_fr* c_s, size_t count, int precision_bits,
                              int lookup_bits, int shift_bits, int a_num_bits, h2svd_fr* out_q,
                              h2svd_fr* out_wit) {
    REQUIRE(ctx && c_s && out_q && out_wit, "rescale_witness: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_rescale(ctx, as_fr(c_s), count, precision_bits, lookup_bits, shift_bits, a_num_bits,
                          as_fr(out_q), as_fr(out_wit));
}

int h2svd_fr_matmul_rescale_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, size_t n, size_t k, size_t m,
                                int precision_bits, int lookup_bits, int shift_bits, int a_num_bits, h2svd_fr* c_s,
                                h2svd_fr* out_q, h2svd_fr* out_wit) {
    REQUIRE(ctx && a && b && c_s && out_q && out_wit, "fr_matmul_rescale: null argument");
    REQUIRE(k >= 1, "fr_matmul_rescale: empty inner dimension");
    REQUIRE(rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr) > 0,
            "fr_matmul_rescale: rescale parameters out of range");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_fr_matmul_rescale(ctx, as_fr(a), as_fr(b), as_fr(c_s), n, k, m, precision_bits, lookup_bits, shift_bits,
                                    a_num_bits, as_fr(out_q), as_fr(out_wit));
}

static int rescale_witness_host(h2svd_ctx* ctx, const h2svd_fr* c_s, size_t count, int precision_bits, int lookup_bits,
                                int shift_bits, int a_num_bits, h2svd_fr* out_q, h2svd_fr* out_wit) {
    REQUIRE(ctx && c_s && out_q && out_wit, "rescale_witness: null argument");
    const int W = rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr);
    REQUIRE(W > 0, "rescale_witness: parameters out of range");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return H2SVD_OK;
    // Chunked, double-buffered: the D2H of chunk i (copy stream) overlaps the kernel of chunk i+1.
    const size_t F = sizeof(Fr);
    const size_t chunk = count < ((size_t)1 << 17) ? count : ((size_t)1 << 17);
    const size_t need = Carver::need(count * F) + Carver::need(count * F) + 2 * Carver::need(chunk * (size_t)W * F);
    H2SVD_TRY(ws_reserve(ctx, need));
    Carver cv(ctx->ws);
    Fr* dcs = cv.take<Fr>(count);
    Fr* dq = cv.take<Fr>(count);
    Fr* dw[2] = {cv.take<Fr>(chunk * (size_t)W), cv.take<Fr>(chunk * (size_t)W)};
    H2SVD_TRY(h2d(ctx, dcs, c_s, count * F));
    H2SVD_TRY(launch_check_canonical(ctx, dcs, count, ctx->d_flag));
    int buf = 0;
    for (size_t off = 0; off < count; off += chunk, buf ^= 1) {
        const size_t cnt = count - off < chunk ? count - off : chunk;
        // buffer `buf` was last drained by the copy issued two chunks ago
        H2SVD_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev[2 + buf], 0));
        H2SVD_TRY(launch_rescale(ctx, dcs + off, cnt, precision_bits, lookup_bits, shift_bits, a_num_bits,
                                 dq + off, dw[buf]));
        H2SVD_CUDA(cudaEventRecord(ctx->ev[buf], ctx->stream));
        H2SVD_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[buf], 0));
        H2SVD_CUDA(cudaMemcpyAsync(as_fr(out_wit) + off * (size_t)W, dw[buf], cnt * (size_t)W * F,
                                   cudaMemcpyDeviceToHost, ctx->copy_stream));
        H2SVD_CUDA(cudaEventRecord(ctx->ev[2 + buf], ctx->copy_stream));
    }
    H2SVD_TRY(d2h(ctx, out_q, dq, count * F));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    return check_flag(ctx, "rescale_witness");
}

int h2svd_rescale_witness(h2svd_ctx* ctx, const h2svd_fr* c_s, size_t count, int precision_bits, int lookup_bits,
                          int shift_bits, int a_num_bits, h2svd_fr* out_q, h2svd_fr* out_wit) {
    REQUIRE(ctx && c_s && out_q && out_wit, "rescale_witness: null argument");
    const int rc = rescale_witness_host(ctx, c_s, count, precision_bits, lookup_bits, shift_bits, a_num_bits, out_q, out_wit);
    if (rc != H2SVD_OK && rc != H2SVD_EINVAL) drain_after_error(ctx);   // nothing may still be copying into the caller's buffers
    return rc;
}

/* ---- range-check witnesses of the SVD verifier's helpers ---- */
static Fr fr_from_u64(const uint64_t* x) {
    Fr f;
    for (int i = 0; i < 4; i++) {
        f.l[2 * i] = (uint32_t)x[i];
        f.l[2 * i + 1] = (uint32_t)(x[i] >> 32);
    }
    return f;
}
int h2svd_abs_less_than_witness_count(const uint64_t bnd[4], int lookup_bits, int with_diff) {
    REQUIRE(bnd != nullptr, "abs_less_than_witness_count: null bound");
    const int w = abs_less_than_params(fr_from_u64(bnd), lookup_bits, with_diff, nullptr, nullptr);
    REQUIRE(w > 0, "abs_less_than_witness_count: parameters out of range");
    return w;
}
int h2svd_abs_less_than_witness_dev(h2svd_ctx* ctx, const h2svd_fr* x, const h2svd_fr* y, size_t count,
                                    const uint64_t bnd[4], int lookup_bits, h2svd_fr* out_wit) {
    REQUIRE(ctx && x && bnd && out_wit, "abs_less_than_witness: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_abs_less_than(ctx, as_fr(x), y ? as_fr(y) : nullptr, count, fr_from_u64(bnd), lookup_bits, as_fr(out_wit));
}
int h2svd_abs_less_than_witness(h2svd_ctx* ctx, const h2svd_fr* x, const h2svd_fr* y, size_t count,
                                const uint64_t bnd[4], int lookup_bits, h2svd_fr* out_wit) {
    REQUIRE(ctx && x && bnd && out_wit, "abs_less_than_witness: null argument");
    const int W = abs_less_than_params(fr_from_u64(bnd), lookup_bits, y != nullptr, nullptr, nullptr);
    REQUIRE(W > 0, "abs_less_than_witness: parameters out of range");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return H2SVD_OK;
    const size_t F = sizeof(Fr);
    H2SVD_TRY(ws_reserve(ctx, 2 * Carver::need(count * F) + Carver::need(count * (size_t)W * F)));
    Carver cv(ctx->ws);
    Fr* dx = cv.take<Fr>(count);
    Fr* dy = cv.take<Fr>(count);
    Fr* dw = cv.take<Fr>(count * (size_t)W);
    H2SVD_TRY(h2d(ctx, dx, x, count * F));
    H2SVD_TRY(launch_check_canonical(ctx, dx, count, ctx->d_flag));
    if (y) {
        H2SVD_TRY(h2d(ctx, dy, y, count * F));
        H2SVD_TRY(launch_check_canonical(ctx, dy, count, ctx->d_flag));
    }
    H2SVD_TRY(launch_abs_less_than(ctx, dx, y ? dy : nullptr, count, fr_from_u64(bnd), lookup_bits, dw));
    H2SVD_TRY(d2h(ctx, out_wit, dw, count * (size_t)W * F));
    return check_flag(ctx, "abs_less_than_witness");
}
int h2svd_range_check_witness_count(int range_bits, int lookup_bits) {
    const int w = range_check_params(range_bits, lookup_bits, nullptr, nullptr);
    REQUIRE(w >= 0, "range_check_witness_count: parameters out of range");
    return w;
}
int h2svd_range_check_witness_dev(h2svd_ctx* ctx, const h2svd_fr* x, size_t count, int range_bits, int lookup_bits,
                                  h2svd_fr* out_wit) {
    REQUIRE(ctx && x && out_wit, "range_check_witness: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_range_check(ctx, as_fr(x), count, range_bits, lookup_bits, as_fr(out_wit));
}
int h2svd_range_check_witness(h2svd_ctx* ctx, const h2svd_fr* x, size_t count, int range_bits, int lookup_bits,
                              h2svd_fr* out_wit) {
    REQUIRE(ctx && x && out_wit, "range_check_witness: null argument");
    const int W = range_check_params(range_bits, lookup_bits, nullptr, nullptr);
    REQUIRE(W >= 0, "range_check_witness: parameters out of range");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    if (count == 0 || W == 0) return H2SVD_OK;
    const size_t F = sizeof(Fr);
    H2SVD_TRY(ws_reserve(ctx, Carver::need(count * F) + Carver::need(count * (size_t)W * F)));
    Carver cv(ctx->ws);
    Fr* dx = cv.take<Fr>(count);
    Fr* dw = cv.take<Fr>(count * (size_t)W);
    H2SVD_TRY(h2d(ctx, dx, x, count * F));
    H2SVD_TRY(launch_check_canonical(ctx, dx, count, ctx->d_flag));
    H2SVD_TRY(launch_range_check(ctx, dx, count, range_bits, lookup_bits, dw));
    H2SVD_TRY(d2h(ctx, out_wit, dw, count * (size_t)W * F));
    return check_flag(ctx, "range_check_witness");
}
int h2svd_mat_times_diag_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t lda,
                             size_t cols_v, h2svd_fr* out) {
    REQUIRE(ctx && a && v && out, "mat_times_diag: null argument");
    REQUIRE(cols_v <= lda, "mat_times_diag: v.len() <= a[0].len()");  // reference :616
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_mat_times_diag(ctx, as_fr(a), as_fr(v), rows, lda, cols_v, as_fr(out));
}
int h2svd_mat_times_diag(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t lda, size_t cols_v,
                         h2svd_fr* out) {
    REQUIRE(ctx && a && v && out, "mat_times_diag: null argument");
    REQUIRE(cols_v <= lda, "mat_times_diag: v.len() <= a[0].len()");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    if (rows == 0 || cols_v == 0) return H2SVD_OK;
    const size_t F = sizeof(Fr);
    H2SVD_TRY(ws_reserve(ctx, Carver::need(rows * lda * F) + Carver::need(cols_v * F) + Carver::need(rows * cols_v * F)));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(rows * lda);
    Fr* dv = cv.take<Fr>(cols_v);
    Fr* dout = cv.take<Fr>(rows * cols_v);
    H2SVD_TRY(h2d(ctx, da, a, rows * lda * F));
    H2SVD_TRY(h2d(ctx, dv, v, cols_v * F));
    H2SVD_TRY(launch_check_canonical(ctx, da, rows * lda, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dv, cols_v, ctx->d_flag));
    H2SVD_TRY(launch_mat_times_diag(ctx, da, dv, rows, lda, cols_v, dout));
    H2SVD_TRY(d2h(ctx, out, dout, rows * cols_v * F));
    return check_flag(ctx, "mat_times_diag");
}

/* ---- the README.md:34-47 sequence: device-pointer form (two streams, graph-capturable) ---- */
namespace h2svd {

// runs `fn` with the handle's launchers pointed at `st`
struct StreamScope {
    h2svd_ctx* ctx;
    cudaStream_t saved;
    StreamScope(h2svd_ctx* c, cudaStream_t st) : ctx(c), saved(c->stream) { c->stream = st; }
    ~StreamScope() { ctx->stream = saved; }
};

// gamma powers (:316-326), the running sums of rows [bv0, bv1) of b . v (:336) and ALL k row totals (b v) on the current stream
static int bv_part(h2svd_ctx* ctx, const Fr* db, const Fr* dg, size_t k, size_t m, size_t bv0, size_t bv1, Fr* dpow,
                   Fr* dpbv_rows, Fr* dbv) {
    H2SVD_TRY(launch_gamma_powers(ctx, dg, m, dpow));
    if (bv0 == 0 && bv1 == k) return launch_mat_vec_prefix(ctx, db, dpow, k, m, 0, dpbv_rows, dbv);
    // a row-sharded caller: every handle needs all k totals (the second operand of a . (b v)) but emits the running-sum
    // witnesses of its own rows only -- the totals are recomputed lazily instead of exchanged (no collective)
    H2SVD_TRY(launch_mat_vec_totals(ctx, db, dpow, k, m, dbv));
    if (bv1 > bv0) H2SVD_TRY(launch_mat_vec_prefix(ctx, db + bv0 * m, dpow, bv1 - bv0, m, 0, dpbv_rows, nullptr));
    return H2SVD_OK;
}

static int mul_witness_dev(h2svd_ctx* ctx, const Fr* da, const Fr* db, const Fr* dg, size_t rows, size_t k, size_t m, int P,
                           int lb, int S, int A, size_t bv0, size_t bv1, Fr* dc, Fr* dq, Fr* dw, Fr* dpow, Fr* dpcv,
                           Fr* dpbv, Fr* dpabv, Fr* ddiff, Fr* dz, Fr* dinv) {
    H2SVD_TRY(ws_reserve(ctx, Carver::need(k * sizeof(Fr)) + 2 * Carver::need(rows * sizeof(Fr))));
    Carver cv(ctx->ws);
    Fr* dbv = cv.take<Fr>(k);
    Fr* dcsv = cv.take<Fr>(rows);
    Fr* dabv = cv.take<Fr>(rows);
    cudaStream_t main = ctx->stream, side = ctx->side_stream;
    cudaEvent_t ev_fork = ctx->ev[4], ev_mid = ctx->ev[5], ev_join = ctx->ev[6];
    if (ctx->tune.step_schedule == 1) {
        // Experiment: the whole of verify_mul runs AFTER the mat-mul, on the side stream, through the low-register
        // mat-vec kernel whose CTAs fit next to the resident rescale CTAs (store-bound: half their issue slots idle).
        H2SVD_TRY(launch_fr_matmul(ctx, da, db, dc, rows, k, m));                                 // :546
        H2SVD_CUDA(cudaEventRecord(ev_mid, main));
        H2SVD_TRY(launch_rescale(ctx, dc, rows * m, P, lb, S, A, dq, dw));                        // :354 (first: takes its SM slots)
        H2SVD_CUDA(cudaStreamWaitEvent(side, ev_mid, 0));
        {
            StreamScope sc(ctx, side);
            const int saved = ctx->tune.matvec_coreside;
            ctx->tune.matvec_coreside = 1;
            int rc = bv_part(ctx, db, dg, k, m, bv0, bv1, dpow, dpbv, dbv);
            if (rc == H2SVD_OK) rc = launch_mat_vec_prefix(ctx, da, dbv, rows, k, 0, dpabv, dabv);     // :337
            if (rc == H2SVD_OK) rc = launch_mat_vec_prefix(ctx, dc, dpow, rows, m, 0, dpcv, dcsv);     // :335
            ctx->tune.matvec_coreside = saved;
            H2SVD_TRY(rc);
            H2SVD_TRY(launch_is_equal(ctx, dcsv, dabv, rows, ddiff, dz, dinv));                   // :339-341
            H2SVD_CUDA(cudaEventRecord(ev_join, side));
        }
        H2SVD_CUDA(cudaStreamWaitEvent(main, ev_join, 0));
        return H2SVD_OK;
    }
    // fork: the C-independent half of verify_mul (integer-pipe mat-vecs) runs under the mat-mul (tensor pipe)
    H2SVD_CUDA(cudaEventRecord(ev_fork, main));
    H2SVD_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    {
        StreamScope sc(ctx, side);
        H2SVD_TRY(bv_part(ctx, db, dg, k, m, bv0, bv1, dpow, dpbv, dbv));
        H2SVD_TRY(launch_mat_vec_prefix(ctx, da, dbv, rows, k, 0, dpabv, dabv));                  // :337
    }
    H2SVD_TRY(launch_fr_matmul(ctx, da, db, dc, rows, k, m));                                     // :546
    H2SVD_CUDA(cudaEventRecord(ev_mid, main));
    H2SVD_CUDA(cudaStreamWaitEvent(side, ev_mid, 0));
    {
        StreamScope sc(ctx, side);   // C.v (integer pipe) next to the rescale kernel (HBM writes)
        H2SVD_TRY(launch_mat_vec_prefix(ctx, dc, dpow, rows, m, 0, dpcv, dcsv));                  // :335
        H2SVD_TRY(launch_is_equal(ctx, dcsv, dabv, rows, ddiff, dz, dinv));                       // :339-341
        H2SVD_CUDA(cudaEventRecord(ev_join, side));
    }
    H2SVD_TRY(launch_rescale(ctx, dc, rows * m, P, lb, S, A, dq, dw));                            // :354
    H2SVD_CUDA(cudaStreamWaitEvent(main, ev_join, 0));
    return H2SVD_OK;
}

static int mul_witness_host(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* gamma, size_t rows,
                            size_t k, size_t m, int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                            size_t bv_row0, size_t bv_row1, h2svd_fr* c_s, h2svd_fr* q, h2svd_fr* wit, h2svd_fr* powers,
                            h2svd_fr* prefix_cv, h2svd_fr* prefix_bv, h2svd_fr* prefix_abv, h2svd_fr* diff,
                            h2svd_fr* is_zero, h2svd_fr* inv, int W) {
    const size_t F = sizeof(Fr);
    // slab height: 128 rows (one row of mat-mul tiles) unless that is more than ~256 MB of rescale witnesses per slab
    size_t slab = ((size_t)256 << 20) / (m * (size_t)W * F);
    if (slab > 128) slab = 128;
    if (slab < 1) slab = 1;
    if (slab > rows) slab = rows;
    const size_t bvn = bv_row1 - bv_row0;
    const size_t need = Carver::need(rows * k * F) + Carver::need(k * m * F) + 3 * Carver::need(rows * m * F) +
                        Carver::need(F) + Carver::need(m * F) + Carver::need((bvn ? bvn : 1) * m * F) +
                        Carver::need(rows * k * F) + Carver::need(k * F) + 5 * Carver::need(rows * F) +
                        2 * Carver::need(slab * m * (size_t)W * F);
    H2SVD_TRY(ws_reserve(ctx, need));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(rows * k);
    Fr* db = cv.take<Fr>(k * m);
    Fr* dc = cv.take<Fr>(rows * m);
    Fr* dq = cv.take<Fr>(rows * m);
    Fr* dpcv = cv.take<Fr>(rows * m);
    Fr* dg = cv.take<Fr>(1);
    Fr* dpow = cv.take<Fr>(m);
    Fr* dpbv = cv.take<Fr>((bvn ? bvn : 1) * m);
    Fr* dpabv = cv.take<Fr>(rows * k);
    Fr* dbv = cv.take<Fr>(k);
    Fr* dcsv = cv.take<Fr>(rows);
    Fr* dabv = cv.take<Fr>(rows);
    Fr* ddiff = cv.take<Fr>(rows);
    Fr* dz = cv.take<Fr>(rows);
    Fr* dinv = cv.take<Fr>(rows);
    Fr* dw[2] = {cv.take<Fr>(slab * m * (size_t)W), cv.take<Fr>(slab * m * (size_t)W)};
    cudaStream_t cs = ctx->stream, xs = ctx->copy_stream;
    auto to_host = [&](void* dst, const void* src, size_t bytes) -> int {
        if (bytes) H2SVD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, xs));
        return H2SVD_OK;
    };
    // inputs: B and gamma first (the C-independent half of verify_mul starts at once), then the first slab of A
    H2SVD_TRY(h2d(ctx, db, b, k * m * F));
    H2SVD_TRY(h2d(ctx, dg, gamma, F));
    H2SVD_TRY(launch_check_canonical(ctx, db, k * m, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dg, 1, ctx->d_flag));
    H2SVD_TRY(bv_part(ctx, db, dg, k, m, bv_row0, bv_row1, dpow, dpbv, dbv));                     // :316-326, :336
    H2SVD_CUDA(cudaEventRecord(ctx->ev[0], cs));
    H2SVD_CUDA(cudaStreamWaitEvent(xs, ctx->ev[0], 0));
    H2SVD_TRY(to_host(powers, dpow, m * F));
    H2SVD_TRY(to_host(prefix_bv, dpbv, bvn * m * F));
    // make the slab-buffer reuse events signalled
    H2SVD_CUDA(cudaEventRecord(ctx->ev[2], xs));
    H2SVD_CUDA(cudaEventRecord(ctx->ev[3], xs));
    // The product is taken in TWO mat-muls -- the first slab alone (its witnesses leave for the host after B + one slab
    // of A), then all remaining rows at once -- so that B is split into byte planes twice, not once per slab, and every
    // MMA tile but the last is a full 128 rows.
    H2SVD_TRY(h2d(ctx, da, a, slab * k * F));
    H2SVD_TRY(launch_check_canonical(ctx, da, slab * k, ctx->d_flag));
    H2SVD_TRY(launch_fr_matmul(ctx, da, db, dc, slab, k, m));                                     // :546
    int buf = 0;
    for (size_t r0 = 0, nr = 0; r0 < rows; r0 += nr, buf ^= 1) {
        nr = rows - r0 < slab ? rows - r0 : slab;
        H2SVD_CUDA(cudaStreamWaitEvent(cs, ctx->ev[2 + buf], 0));  // slab buffer `buf` drained two slabs ago
        H2SVD_TRY(launch_rescale(ctx, dc + r0 * m, nr * m, precision_bits, lookup_bits, shift_bits, a_num_bits,
                                 dq + r0 * m, dw[buf]));                                          // :354
        H2SVD_TRY(launch_mat_vec_prefix(ctx, dc + r0 * m, dpow, nr, m, 0, dpcv + r0 * m, dcsv + r0));    // :335
        H2SVD_TRY(launch_mat_vec_prefix(ctx, da + r0 * k, dbv, nr, k, 0, dpabv + r0 * k, dabv + r0));    // :337
        H2SVD_TRY(launch_is_equal(ctx, dcsv + r0, dabv + r0, nr, ddiff + r0, dz + r0, dinv + r0));       // :339-341
        H2SVD_CUDA(cudaEventRecord(ctx->ev[buf], cs));
        if (r0 == 0 && rows > slab) {
            // the rest of A and its product, queued behind the first slab's kernels and under its D2H
            H2SVD_TRY(h2d(ctx, da + slab * k, as_fr(a) + slab * k, (rows - slab) * k * F));
            H2SVD_TRY(launch_check_canonical(ctx, da + slab * k, (rows - slab) * k, ctx->d_flag));
            H2SVD_TRY(launch_fr_matmul(ctx, da + slab * k, db, dc + slab * m, rows - slab, k, m));
        }
        H2SVD_CUDA(cudaStreamWaitEvent(xs, ctx->ev[buf], 0));
        H2SVD_TRY(to_host(as_fr(wit) + r0 * m * (size_t)W, dw[buf], nr * m * (size_t)W * F));
        H2SVD_CUDA(cudaEventRecord(ctx->ev[2 + buf], xs));
        H2SVD_TRY(to_host(as_fr(c_s) + r0 * m, dc + r0 * m, nr * m * F));
        H2SVD_TRY(to_host(as_fr(q) + r0 * m, dq + r0 * m, nr * m * F));
        H2SVD_TRY(to_host(as_fr(prefix_cv) + r0 * m, dpcv + r0 * m, nr * m * F));
        H2SVD_TRY(to_host(as_fr(prefix_abv) + r0 * k, dpabv + r0 * k, nr * k * F));
    }
    H2SVD_CUDA(cudaEventRecord(ctx->ev[0], cs));
    H2SVD_CUDA(cudaStreamWaitEvent(xs, ctx->ev[0], 0));
    H2SVD_TRY(to_host(diff, ddiff, rows * F));
    H2SVD_TRY(to_host(is_zero, dz, rows * F));
    H2SVD_TRY(to_host(inv, dinv, rows * F));
    H2SVD_CUDA(cudaStreamSynchronize(xs));
    return check_flag(ctx, "zkmatrix_mul_witness");
}

}  // namespace h2svd

int h2svd_zkmatrix_mul_witness_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* gamma, size_t rows,
                                   size_t k, size_t m, int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                                   size_t bv_row0, size_t bv_row1, h2svd_fr* c_s, h2svd_fr* q, h2svd_fr* wit,
                                   h2svd_fr* powers, h2svd_fr* prefix_cv, h2svd_fr* prefix_bv, h2svd_fr* prefix_abv,
                                   h2svd_fr* diff, h2svd_fr* is_zero, h2svd_fr* inv) {
    REQUIRE(ctx && a && b && gamma && c_s && q && wit && powers && prefix_cv && prefix_abv && diff && is_zero && inv,
            "zkmatrix_mul_witness_dev: null argument");
    REQUIRE(rows >= 1 && k >= 1 && m >= 1, "zkmatrix_mul_witness_dev: empty matrix");
    REQUIRE(bv_row0 <= bv_row1 && bv_row1 <= k, "zkmatrix_mul_witness_dev: bad prefix_bv row range");
    REQUIRE(bv_row0 == bv_row1 || prefix_bv, "zkmatrix_mul_witness_dev: null prefix_bv");
    REQUIRE(rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr) > 0,
            "zkmatrix_mul_witness_dev: rescale parameters out of range");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return mul_witness_dev(ctx, as_fr(a), as_fr(b), as_fr(gamma), rows, k, m, precision_bits, lookup_bits, shift_bits,
                           a_num_bits, bv_row0, bv_row1, as_fr(c_s), as_fr(q), as_fr(wit), as_fr(powers), as_fr(prefix_cv),
                           prefix_bv ? as_fr(prefix_bv) : nullptr, as_fr(prefix_abv), as_fr(diff), as_fr(is_zero),
                           as_fr(inv));
}

int h2svd_mat_vec_totals_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* v, size_t rows, size_t len,
                             h2svd_fr* out_totals) {
    REQUIRE(ctx && a && v && out_totals, "mat_vec_totals: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_mat_vec_totals(ctx, as_fr(a), as_fr(v), rows, len, as_fr(out_totals));
}

/* ---- fused, slab-pipelined sequence (host pointers) ---- */
int h2svd_zkmatrix_mul_witness(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* gamma, size_t rows,
                               size_t k, size_t m, int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                               size_t bv_row0, size_t bv_row1, h2svd_fr* c_s, h2svd_fr* q, h2svd_fr* wit,
                               h2svd_fr* powers, h2svd_fr* prefix_cv, h2svd_fr* prefix_bv, h2svd_fr* prefix_abv,
                               h2svd_fr* diff, h2svd_fr* is_zero, h2svd_fr* inv) {
    REQUIRE(ctx && a && b && gamma && c_s && q && wit && powers && prefix_cv && prefix_abv && diff && is_zero && inv,
            "zkmatrix_mul_witness: null argument");
    REQUIRE(rows >= 1 && k >= 1 && m >= 1, "zkmatrix_mul_witness: empty matrix");
    REQUIRE(bv_row0 <= bv_row1 && bv_row1 <= k, "zkmatrix_mul_witness: bad prefix_bv row range");
    REQUIRE(bv_row0 == bv_row1 || prefix_bv, "zkmatrix_mul_witness: null prefix_bv");
    const int W = rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr);
    REQUIRE(W > 0, "zkmatrix_mul_witness: rescale parameters out of range");
    REQUIRE(!ctx->capturing, "zkmatrix_mul_witness: host-pointer entry points cannot be captured into a graph");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const int rc = mul_witness_host(ctx, a, b, gamma, rows, k, m, precision_bits, lookup_bits, shift_bits, a_num_bits,
                                    bv_row0, bv_row1, c_s, q, wit, powers, prefix_cv, prefix_bv, prefix_abv, diff, is_zero,
                                    inv, W);
    if (rc != H2SVD_OK) drain_after_error(ctx);
    return rc;
}

/* ---- CUDA-graph capture of any sequence of *_dev calls on one handle ---- */
struct h2svd_graph {
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;        // kernel launches recorded between begin and end
    uint64_t ws_generation = 0;   // the handle's workspace generation the recorded pointers belong to
};

int h2svd_graph_begin(h2svd_ctx* ctx) {
    REQUIRE(ctx, "graph_begin: null handle");
    REQUIRE(!ctx->capturing, "graph_begin: a capture is already in progress on this handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    H2SVD_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->capture_launches0 = ctx->launches;
    return H2SVD_OK;
}

int h2svd_graph_end(h2svd_ctx* ctx, h2svd_graph** out) {
    REQUIRE(ctx && out, "graph_end: null argument");
    REQUIRE(ctx->capturing, "graph_end: no capture in progress");
    *out = nullptr;
    ctx->capturing = false;
    cudaGraph_t g = nullptr;
    H2SVD_CUDA(cudaStreamEndCapture(ctx->stream, &g));
    h2svd_graph* gr = new (std::nothrow) h2svd_graph();
    if (!gr) {
        cudaGraphDestroy(g);
        return H2SVD_ENOMEM;
    }
    const cudaError_t e = cudaGraphInstantiate(&gr->exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        delete gr;
        return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
    }
    gr->launches = ctx->launches - ctx->capture_launches0;
    gr->ws_generation = ctx->ws_generation;
    ctx->launches = ctx->capture_launches0;   // nothing has run yet: h2svd_graph_launch accounts for every replay
    *out = gr;
    return H2SVD_OK;
}

int h2svd_graph_launch(h2svd_ctx* ctx, h2svd_graph* graph) {
    REQUIRE(ctx && graph && graph->exec, "graph_launch: null argument");
    REQUIRE(graph->ws_generation == ctx->ws_generation,
            "graph_launch: a workspace of the handle was reallocated after this graph was recorded (a larger call ran since): "
            "record it again");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    H2SVD_CUDA(cudaGraphLaunch(graph->exec, ctx->stream));
    ctx->launches += graph->launches;
    return H2SVD_OK;
}

void h2svd_graph_destroy(h2svd_graph* graph) {
    if (!graph) return;
    if (graph->exec) cudaGraphExecDestroy(graph->exec);
    delete graph;
}

/* ---- one process, several GPUs: rows of A / C partitioned over the handles, B replicated ---- */
struct h2svd_multi {
    std::vector<h2svd_ctx*> ctx;
};

int h2svd_multi_create(h2svd_multi** out, const int* devices, int n_dev) {
    REQUIRE(out && devices, "multi_create: null argument");
    REQUIRE(n_dev >= 1 && n_dev <= 64, "multi_create: device count out of range");
    *out = nullptr;
    h2svd_multi* mh = new (std::nothrow) h2svd_multi();
    if (!mh) return H2SVD_ENOMEM;
    for (int i = 0; i < n_dev; i++) {
        h2svd_ctx* c = nullptr;
        const int rc = h2svd_create(&c, devices[i], nullptr);   // the same device may appear more than once
        if (rc != H2SVD_OK) {
            h2svd_multi_destroy(mh);
            return rc;
        }
        mh->ctx.push_back(c);
    }
    *out = mh;
    return H2SVD_OK;
}

void h2svd_multi_destroy(h2svd_multi* mh) {
    if (!mh) return;
    for (h2svd_ctx* c : mh->ctx) h2svd_destroy(c);
    delete mh;
}

int h2svd_multi_count(h2svd_multi* mh) { return mh ? (int)mh->ctx.size() : 0; }
h2svd_ctx* h2svd_multi_ctx(h2svd_multi* mh, int i) { return (mh && i >= 0 && i < (int)mh->ctx.size()) ? mh->ctx[i] : nullptr; }

// contiguous near-equal split: the first (total % parts) parts get one extra row
static void split_rows(size_t total, size_t parts, size_t idx, size_t* lo, size_t* hi) {
    const size_t base = total / parts, extra = total % parts;
    *lo = idx * base + (idx < extra ? idx : extra);
    *hi = *lo + base + (idx < extra ? 1 : 0);
}

int h2svd_multi_zkmatrix_mul_witness(h2svd_multi* mh, const h2svd_fr* a, const h2svd_fr* b, const h2svd_fr* gamma, size_t n,
                                     size_t k, size_t m, int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                                     h2svd_fr* c_s, h2svd_fr* q, h2svd_fr* wit, h2svd_fr* powers, h2svd_fr* prefix_cv,
                                     h2svd_fr* prefix_bv, h2svd_fr* prefix_abv, h2svd_fr* diff, h2svd_fr* is_zero,
                                     h2svd_fr* inv) {
    REQUIRE(mh && !mh->ctx.empty(), "multi_zkmatrix_mul_witness: null handle");
    REQUIRE(a && b && gamma && c_s && q && wit && powers && prefix_cv && prefix_bv && prefix_abv && diff && is_zero && inv,
            "multi_zkmatrix_mul_witness: null argument");
    REQUIRE(n >= 1 && k >= 1 && m >= 1, "multi_zkmatrix_mul_witness: empty matrix");
    const int W = rescale_params(precision_bits, lookup_bits, shift_bits, a_num_bits, nullptr, nullptr);
    REQUIRE(W > 0, "multi_zkmatrix_mul_witness: rescale parameters out of range");
    // every participating handle gets at least one row of A; the rows of b . v are split over the same handles.  No
    // exchange step: each handle derives all k totals of (b v) itself (h2svd_zkmatrix_mul_witness), field addition is
    // exact, so the assembled witness is byte-identical to the single-GPU one.
    const size_t parts = mh->ctx.size() < n ? mh->ctx.size() : n;
    std::vector<int> rc(parts, H2SVD_OK);
    std::vector<std::string> msg(parts);
    std::vector<std::vector<h2svd_fr>> pow_tmp(parts);
    std::vector<std::thread> th;
    for (size_t g = 0; g < parts; g++) {
        th.emplace_back([&, g]() {
            size_t r0, r1, b0, b1;
            split_rows(n, parts, g, &r0, &r1);
            split_rows(k, parts, g, &b0, &b1);
            h2svd_fr* pw = powers;
            if (g != 0) {   // every handle returns the same gamma powers: only the first writes the caller's buffer
                pow_tmp[g].resize(m);
                pw = pow_tmp[g].data();
            }
            rc[g] = h2svd_zkmatrix_mul_witness(mh->ctx[g], a + r0 * k, b, gamma, r1 - r0, k, m, precision_bits, lookup_bits,
                                               shift_bits, a_num_bits, b0, b1, c_s + r0 * m, q + r0 * m,
                                               wit + r0 * m * (size_t)W, pw, prefix_cv + r0 * m, prefix_bv + b0 * m,
                                               prefix_abv + r0 * k, diff + r0, is_zero + r0, inv + r0);
            if (rc[g] != H2SVD_OK) msg[g] = h2svd_last_error();   // the message is thread-local: carry it over
        });
    }
    for (auto& t : th) t.join();
    for (size_t g = 0; g < parts; g++)
        if (rc[g] != H2SVD_OK) {
            set_error("device %d (part %zu of %zu): %s", mh->ctx[g]->device, g, parts, msg[g].c_str());
            return rc[g];
        }
    return H2SVD_OK;
}

int h2svd_host_alloc(size_t bytes, void** out) {
    REQUIRE(out != nullptr, "host_alloc: out is null");
    *out = nullptr;
    H2SVD_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return H2SVD_OK;
}
void h2svd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

/* ---- K5/K6 ---- */
int h2svd_zkvec_inner_prefix_dev(h2svd_ctx* ctx, const h2svd_fr* x, const h2svd_fr* self, size_t batch,
                                 size_t len, h2svd_fr* out_prefix) {
    REQUIRE(ctx && x && self && out_prefix, "zkvec_inner_prefix: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    // gate.inner_product(u = x, v = self): reference src/matrix/mod.rs:100
    return launch_mat_vec_prefix(ctx, as_fr(x), as_fr(self), batch, len, len, as_fr(out_prefix), nullptr);
}
int h2svd_zkvec_inner_prefix(h2svd_ctx* ctx, const h2svd_fr* x, const h2svd_fr* self, size_t batch, size_t len,
                             h2svd_fr* out_prefix) {
    REQUIRE(ctx && x && self && out_prefix, "zkvec_inner_prefix: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t cnt = batch * len, F = sizeof(Fr);
    if (cnt == 0) return H2SVD_OK;
    H2SVD_TRY(ws_reserve(ctx, 3 * Carver::need(cnt * F)));
    Carver cv(ctx->ws);
    Fr* dx = cv.take<Fr>(cnt);
    Fr* ds = cv.take<Fr>(cnt);
    Fr* dout = cv.take<Fr>(cnt);
    H2SVD_TRY(h2d(ctx, dx, x, cnt * F));
    H2SVD_TRY(h2d(ctx, ds, self, cnt * F));
    H2SVD_TRY(launch_check_canonical(ctx, dx, cnt, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, ds, cnt, ctx->d_flag));
    H2SVD_TRY(launch_mat_vec_prefix(ctx, dx, ds, batch, len, len, dout, nullptr));
    H2SVD_TRY(d2h(ctx, out_prefix, dout, cnt * F));
    return check_flag(ctx, "zkvec_inner_prefix");
}
int h2svd_zkvec_sub_dev(h2svd_ctx* ctx, const h2svd_fr* self, const h2svd_fr* x, size_t count, h2svd_fr* out) {
    REQUIRE(ctx && self && x && out, "zkvec_sub: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_sub(ctx, as_fr(self), as_fr(x), count, as_fr(out));
}
int h2svd_zkvec_sub(h2svd_ctx* ctx, const h2svd_fr* self, const h2svd_fr* x, size_t count, h2svd_fr* out) {
    REQUIRE(ctx && self && x && out, "zkvec_sub: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t F = sizeof(Fr);
    if (count == 0) return H2SVD_OK;
    H2SVD_TRY(ws_reserve(ctx, 3 * Carver::need(count * F)));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(count);
    Fr* db = cv.take<Fr>(count);
    Fr* dout = cv.take<Fr>(count);
    H2SVD_TRY(h2d(ctx, da, self, count * F));
    H2SVD_TRY(h2d(ctx, db, x, count * F));
    H2SVD_TRY(launch_check_canonical(ctx, da, count, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, db, count, ctx->d_flag));
    H2SVD_TRY(launch_sub(ctx, da, db, count, dout));
    H2SVD_TRY(d2h(ctx, out, dout, count * F));
    return check_flag(ctx, "zkvec_sub");
}
int h2svd_isqrt_fixed_dev(h2svd_ctx* ctx, const h2svd_fr* a, size_t count, int precision_bits, h2svd_fr* out) {
    REQUIRE(ctx && a && out, "isqrt_fixed: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_isqrt(ctx, as_fr(a), count, precision_bits, as_fr(out));
}
int h2svd_isqrt_fixed(h2svd_ctx* ctx, const h2svd_fr* a, size_t count, int precision_bits, h2svd_fr* out) {
    REQUIRE(ctx && a && out, "isqrt_fixed: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t F = sizeof(Fr);
    if (count == 0) return H2SVD_OK;
    H2SVD_TRY(ws_reserve(ctx, 2 * Carver::need(count * F)));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(count);
    Fr* dout = cv.take<Fr>(count);
    H2SVD_TRY(h2d(ctx, da, a, count * F));
    H2SVD_TRY(launch_check_canonical(ctx, da, count, ctx->d_flag));
    H2SVD_TRY(launch_isqrt(ctx, da, count, precision_bits, dout));
    H2SVD_TRY(d2h(ctx, out, dout, count * F));
    return check_flag(ctx, "isqrt_fixed");
}

/* ---- quantization ---- */
int h2svd_quantize_dev(h2svd_ctx* ctx, const double* x, size_t count, int precision_bits, h2svd_fr* out) {
    REQUIRE(ctx && x && out, "quantize: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_quantize(ctx, x, count, precision_bits, as_fr(out));
}
int h2svd_quantize(h2svd_ctx* ctx, const double* x, size_t count, int precision_bits, h2svd_fr* out) {
    REQUIRE(ctx && x && out, "quantize: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    if (count == 0) return H2SVD_OK;
    H2SVD_TRY(ws_reserve(ctx, Carver::need(count * sizeof(double)) + Carver::need(count * sizeof(Fr))));
    Carver cv(ctx->ws);
    double* dx = cv.take<double>(count);
    Fr* dout = cv.take<Fr>(count);
    H2SVD_TRY(h2d(ctx, dx, x, count * sizeof(double)));
    H2SVD_TRY(launch_quantize(ctx, dx, count, precision_bits, dout));
    H2SVD_TRY(d2h(ctx, out, dout, count * sizeof(Fr)));
    return check_flag(ctx, "quantize");
}

int h2svd_check_canonical_dev(h2svd_ctx* ctx, const h2svd_fr* x, size_t count) {
    REQUIRE(ctx && x, "check_canonical: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    H2SVD_TRY(launch_check_canonical(ctx, as_fr(x), count, ctx->d_flag));
    return check_flag(ctx, "check_canonical");
}

/* ---- host-side scalar helpers (fr.cuh is host-callable) ---- */
static void fr_to_u64(const Fr& f, uint64_t* x) {
    for (int i = 0; i < 4; i++) x[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32);
}
int h2svd_host_fr_from_canonical(const uint64_t x[4], h2svd_fr* out) {
    REQUIRE(x && out, "host_fr_from_canonical: null argument");
    const Fr v = fr_from_u64(x);
    if (!fr::is_canonical(v)) {
        set_error("host_fr_from_canonical: value >= r");
        return H2SVD_ERANGE;
    }
    fr_to_u64(fr::to_mont(v), out->l);
    return H2SVD_OK;
}
void h2svd_host_fr_to_canonical(const h2svd_fr* a, uint64_t out[4]) { fr_to_u64(fr::from_mont(fr_from_u64(a->l)), out); }
void h2svd_host_fr_add(const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* out) {
    fr_to_u64(fr::add(fr_from_u64(a->l), fr_from_u64(b->l)), out->l);
}
void h2svd_host_fr_sub(const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* out) {
    fr_to_u64(fr::sub(fr_from_u64(a->l), fr_from_u64(b->l)), out->l);
}
void h2svd_host_fr_mul(const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* out) {
    fr_to_u64(fr::mont_mul(fr_from_u64(a->l), fr_from_u64(b->l)), out->l);
}

int h2svd_microbench_imad(h2svd_ctx* ctx, int kind, int iters, double* ops_per_s) {
    REQUIRE(ctx, "microbench: null handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_microbench(ctx, kind, iters, ops_per_s);
}

int h2svd_microbench_hbm(h2svd_ctx* ctx, int kind, size_t bytes, double* gb_per_s) {
    REQUIRE(ctx, "microbench_hbm: null handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_microbench_hbm(ctx, kind, bytes, gb_per_s);
}

int h2svd_microbench_tensor_i8(h2svd_ctx* ctx, int kind, double min_seconds, double* ops_per_s) {
    REQUIRE(ctx, "microbench_tensor_i8: null handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_microbench_i8(ctx, kind, min_seconds, ops_per_s);
}

/* debug / triage only (not in the public header): per-handle tuning switches and the engine of the last mat-mul */
int h2svd_debug_tune(h2svd_ctx* ctx, const char* key, int value) {
    REQUIRE(ctx && key, "debug_tune: null argument");
    struct { const char* name; int* slot; } const keys[] = {
        {"matmul_tc", &ctx->tune.matmul_tc},       {"matmul_small", &ctx->tune.matmul_small},
        {"matmul_small_width", &ctx->tune.matmul_small_width},
        {"matmul_cluster", &ctx->tune.matmul_cluster},
        {"matmul_karatsuba", &ctx->tune.kara},     {"matmul_streamk", &ctx->tune.streamk},
        {"matmul_variant", &ctx->tune.variant},    {"fuse_rescale", &ctx->tune.fuse_rescale},
        {"rescale_generic", &ctx->tune.rescale_generic}, {"matvec_warp_kernel", &ctx->tune.matvec_warp},
        {"matvec_seg", &ctx->tune.matvec_seg},       {"matvec_x2", &ctx->tune.matvec_x2},
        {"matvec_segs", &ctx->tune.matvec_segs},       {"rescale_ch", &ctx->tune.rescale_ch},
        {"rescale_store", &ctx->tune.rescale_store}, {"rescale_fast_sums", &ctx->tune.rescale_fast_sums},
        {"rescale_ctas", &ctx->tune.rescale_ctas}, {"matvec_coreside", &ctx->tune.matvec_coreside}, {"step_schedule", &ctx->tune.step_schedule},
    };
    for (const auto& e : keys)
        if (strcmp(e.name, key) == 0) {
            *e.slot = value;
            return H2SVD_OK;
        }
    set_error("debug_tune: unknown key '%s'", key);
    return H2SVD_EINVAL;
}
/* enable != 0: allocate (and zero) the 128-slot timeline the tensor-core mat-mul kernels stamp; out (may be null): copy it
 * to the host after synchronising.  enable == 0: switch it off again. */
int h2svd_debug_matmul_timeline(h2svd_ctx* ctx, int enable, unsigned long long* out) {
    REQUIRE(ctx, "debug_matmul_timeline: null handle");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out && ctx->d_timeline) H2SVD_CUDA(cudaMemcpy(out, ctx->d_timeline, 128 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (enable) {
        if (!ctx->d_timeline) H2SVD_CUDA(cudaMalloc((void**)&ctx->d_timeline, 128 * sizeof(unsigned long long)));
        H2SVD_CUDA(cudaMemset(ctx->d_timeline, 0, 128 * sizeof(unsigned long long)));
    } else if (ctx->d_timeline) {
        cudaFree(ctx->d_timeline);
        ctx->d_timeline = nullptr;
    }
    return H2SVD_OK;
}

int h2svd_debug_last_matmul_engine(h2svd_ctx* ctx) {
    if (!ctx) return -1;
    if (ctx->last_engine != 3) return ctx->last_engine;
    // both tensor-core engines were enqueued: the device flag says which one did the work (synchronises the stream)
    int mode = 1;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
        cudaMemcpy(&mode, ctx->d_mode, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    return mode == 0 ? 3 : 2;
}

/* debug / triage only: fully reduced one-thread-per-element mat-mul on device pointers */
int h2svd_debug_fr_matmul_naive_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* c, size_t n,
                                    size_t k, size_t m) {
    REQUIRE(ctx && a && b && c, "fr_matmul_naive: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_fr_matmul_naive(ctx, as_fr(a), as_fr(b), as_fr(c), n, k, m);
}

}  // extern "C"
