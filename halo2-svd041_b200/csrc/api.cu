// C ABI of libh2svd_b200 (include/h2svd_b200.h): handle management, argument checking, host<->device
// staging for the host-pointer entry points.  No CPU fallback anywhere: every entry point either runs
// the CUDA kernels or fails with a negative code.
#include <stdarg.h>
#include <string.h>

#include <new>

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
    // the old block may still be in use by queued work: drain first
    H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
    H2SVD_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->ws) H2SVD_CUDA(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    H2SVD_CUDA(cudaMalloc(&ctx->ws, want));
    ctx->ws_bytes = want;
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
        if (flag == 3) {  // a bounded pipeline wait of the tensor-core mat-mul expired (it traps: normally unreachable)
            set_error("%s: internal error: tensor-core mat-mul pipeline timed out", what);
            return H2SVD_ECUDA;
        }
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
        cudaMalloc((void**)&ctx->d_flag, sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_flag, 0, sizeof(int)) != cudaSuccess) {
        h2svd_destroy(ctx);
        return cuda_fail(cudaGetLastError(), "handle setup", __FILE__, __LINE__);
    }
    for (int i = 0; i < 4; i++) {
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
    for (int i = 0; i < 4; i++)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->sk_ws) cudaFree(ctx->sk_ws);
    if (ctx->kara_ws) cudaFree(ctx->kara_ws);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
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

int h2svd_rescale_witness(h2svd_ctx* ctx, const h2svd_fr* c_s, size_t count, int precision_bits, int lookup_bits,
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

/* ---- fused, slab-pipelined sequence ---- */
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
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    const size_t F = sizeof(Fr);
    // slab height: ~96 MB of rescale witnesses per slab (two slab buffers are in flight)
    size_t slab = ((size_t)96 << 20) / (m * (size_t)W * F);
    if (slab < 1) slab = 1;
    if (slab > rows) slab = rows;
    const size_t need = Carver::need(rows * k * F) + Carver::need(k * m * F) + 3 * Carver::need(rows * m * F) +
                        Carver::need(F) + Carver::need(m * F) + Carver::need(k * m * F) + Carver::need(rows * k * F) +
                        Carver::need(k * F) + 5 * Carver::need(rows * F) + 2 * Carver::need(slab * m * (size_t)W * F);
    H2SVD_TRY(ws_reserve(ctx, need));
    Carver cv(ctx->ws);
    Fr* da = cv.take<Fr>(rows * k);
    Fr* db = cv.take<Fr>(k * m);
    Fr* dc = cv.take<Fr>(rows * m);
    Fr* dq = cv.take<Fr>(rows * m);
    Fr* dpcv = cv.take<Fr>(rows * m);
    Fr* dg = cv.take<Fr>(1);
    Fr* dpow = cv.take<Fr>(m);
    Fr* dpbv = cv.take<Fr>(k * m);
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
    // inputs: B and gamma first (the C-independent half of verify_mul starts at once); A follows slab by slab
    H2SVD_TRY(h2d(ctx, db, b, k * m * F));
    H2SVD_TRY(h2d(ctx, dg, gamma, F));
    H2SVD_TRY(launch_check_canonical(ctx, db, k * m, ctx->d_flag));
    H2SVD_TRY(launch_check_canonical(ctx, dg, 1, ctx->d_flag));
    H2SVD_TRY(launch_gamma_powers(ctx, dg, m, dpow));                               // :316-326
    H2SVD_TRY(launch_mat_vec_prefix(ctx, db, dpow, k, m, 0, dpbv, dbv));            // :336
    H2SVD_CUDA(cudaEventRecord(ctx->ev[0], cs));
    H2SVD_CUDA(cudaStreamWaitEvent(xs, ctx->ev[0], 0));
    H2SVD_TRY(to_host(powers, dpow, m * F));
    H2SVD_TRY(to_host(prefix_bv, dpbv + bv_row0 * m, (bv_row1 - bv_row0) * m * F));
    // the copy stream must not start overwriting dw[] users... nothing pending yet; make the reuse events signalled
    H2SVD_CUDA(cudaEventRecord(ctx->ev[2], xs));
    H2SVD_CUDA(cudaEventRecord(ctx->ev[3], xs));
    int buf = 0;
    for (size_t r0 = 0, nr = 0; r0 < rows; r0 += nr, buf ^= 1) {
        // a short first slab gets the copy engine going early; after that the copies are the bottleneck anyway
        const size_t want = r0 == 0 && slab >= 8 ? slab / 8 : slab;
        nr = rows - r0 < want ? rows - r0 : want;
        // A is uploaded slab by slab too: the first witnesses leave for the host after B + one slab of A, not after all of A
        H2SVD_TRY(h2d(ctx, da + r0 * k, as_fr(a) + r0 * k, nr * k * F));
        H2SVD_TRY(launch_check_canonical(ctx, da + r0 * k, nr * k, ctx->d_flag));
        H2SVD_CUDA(cudaStreamWaitEvent(cs, ctx->ev[2 + buf], 0));  // slab buffer `buf` drained two slabs ago
        H2SVD_TRY(launch_fr_matmul_rescale(ctx, da + r0 * k, db, dc + r0 * m, nr, k, m, precision_bits, lookup_bits,
                                           shift_bits, a_num_bits, dq + r0 * m, dw[buf]));               // :546, :354
        H2SVD_TRY(launch_mat_vec_prefix(ctx, dc + r0 * m, dpow, nr, m, 0, dpcv + r0 * m, dcsv + r0));    // :335
        H2SVD_TRY(launch_mat_vec_prefix(ctx, da + r0 * k, dbv, nr, k, 0, dpabv + r0 * k, dabv + r0));    // :337
        H2SVD_TRY(launch_is_equal(ctx, dcsv + r0, dabv + r0, nr, ddiff + r0, dz + r0, dinv + r0));       // :339-341
        H2SVD_CUDA(cudaEventRecord(ctx->ev[buf], cs));
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

/* debug / triage only: fully reduced one-thread-per-element mat-mul on device pointers */
int h2svd_debug_fr_matmul_naive_dev(h2svd_ctx* ctx, const h2svd_fr* a, const h2svd_fr* b, h2svd_fr* c, size_t n,
                                    size_t k, size_t m) {
    REQUIRE(ctx && a && b && c, "fr_matmul_naive: null argument");
    H2SVD_CUDA(cudaSetDevice(ctx->device));
    return launch_fr_matmul_naive(ctx, as_fr(a), as_fr(b), as_fr(c), n, k, m);
}

}  // extern "C"
