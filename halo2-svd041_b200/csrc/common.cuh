// Shared host/device plumbing of libh2svd_b200: handle, error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/h2svd_b200.h"
#include "fr.cuh"

struct h2svd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int sm_count = 0;
    // grow-only device workspace (scratch for transposes, host-pointer entry points, ...)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // second stream + events for copy/compute overlap in the host-pointer entry points
    cudaStream_t copy_stream = nullptr;
    // third stream: the C-independent half of verify_mul (integer-pipe mat-vecs) under the mat-mul, C.v next to the rescale
    cudaStream_t side_stream = nullptr;
    bool capturing = false;       // between h2svd_graph_begin and h2svd_graph_end
    uint64_t capture_launches0 = 0;
    uint64_t ws_generation = 0;   // bumped whenever a workspace is reallocated: recorded graphs hold the old pointers
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 0-3 copy pipelines, 4-6 fork/mid/join of the two-stream step
    void* kara_ws = nullptr;  // pre-split (Karatsuba) operands of the mat-mul
    size_t kara_ws_bytes = 0;
    void* sk_ws = nullptr;  // stream-K partial-tile workspace of the mat-mul
    size_t sk_ws_bytes = 0;
    int* d_flag = nullptr;  // device flag for validation kernels
    int* d_mode = nullptr;  // device flag of the tensor-core mat-mul: 0 = small-operand engine, 1 = full-width engine (d_flag + 1)
    uint64_t launches = 0;
    unsigned long long* d_timeline = nullptr;  // triage: phase timestamps of the tensor-core mat-mul (h2svd_debug_matmul_timeline)
    // Triage / tuning switches (h2svd_debug_tune).  Per handle: two handles on two streams never see each other's settings.
    struct Tuning {
        int matmul_tc = -1;       // tensor-core mat-mul engines: -1 auto, 0 never, 1 always
        int matmul_small = -1;    // small-operand tensor-core engine: -1 auto (range-detected on the device), 0 never
        int matmul_cluster = 0;      // small-operand engine: clusters of 2 CTAs share each A-plane stage by TMA multicast (0: off; measured: no gain)
        int matmul_small_width = 0;  // tile width of the small-operand engine: 0 auto (cost model), 8 / 16 / 24 / 28 forced
        int kara = -1;            // IMAD engines: -1 auto, 0 schoolbook kernels only, 1..3 force a Karatsuba variant
        int streamk = -1;         // IMAD engines: -1 auto, 0 never, 1 always use the stream-K schedule
        int variant = 0;          // schoolbook tile variant
        int fuse_rescale = 0;     // 1: rescale witnesses from the tensor-core epilogue (experimental)
        int rescale_generic = 0;  // 1: force the generic (unstaged) rescale kernel
        int rescale_store = 0;    // witness stream of the rescale kernel: 0 auto (1 when the layout allows, else 3), 1 256-byte-aligned TMA tensor stores, 2 coalesced STG, 3 per-row bulk copies
        int matvec_coreside = 0;     // mat-vecs through the low-register, shared-memory-free kernel (fits next to the rescale CTAs)
        int step_schedule = 0;       // zkmatrix_mul_witness_dev: 1 = every Freivalds mat-vec AFTER the mat-mul, co-resident with the rescale kernel
        int rescale_ctas = 3;        // resident CTAs per SM of the TMA-store rescale kernel (2 or 3)
        int rescale_fast_sums = 0;   // 1: range-check running sums through fr::SmallSum (14 % fewer instructions, measured 1 % SLOWER: the kernel is store-bound)
        int rescale_ch = 8;       // witnesses per bulk store of the staged rescale kernel: 4, 6 or 8
        int matvec_warp = 0;      // 1: force the warp-per-segment mat-vec prefix kernel
        int matvec_x2 = 1;        // mat-vec prefix kernels: 1 = two Montgomery products at a time with interleaved carry chains
        int matvec_segs = 0;      // warps per row of the several-warps-per-row kernel: 0 auto (by row length), 2 / 4 / 8
        int matvec_seg = -1;      // several-warps-per-row mat-vec prefix kernel: -1 auto (few long rows), 0 never, 1 always
    } tune;
    int last_engine = -1;         // engine of the last mat-mul launch: 0 schoolbook, 1 Karatsuba, 2 tensor core, 3 small-operand
};

namespace h2svd {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
// Ensures ctx->ws has at least `bytes`; returns H2SVD_OK / H2SVD_ENOMEM.
int ws_reserve(h2svd_ctx* ctx, size_t bytes);
// Same for a dedicated buffer (*buf, *cur bytes): grows it to `need` (drains the streams first; refused during graph capture).
int ws_grow(h2svd_ctx* ctx, void** buf, size_t* cur, size_t need);

#define H2SVD_CUDA(call)                                                        \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return h2svd::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define H2SVD_LAUNCH_CHECK(ctx)                 \
    do {                                        \
        (ctx)->launches++;                      \
        H2SVD_CUDA(cudaGetLastError());         \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per function AND per device: remember it per device so that
// several handles on different GPUs in one process each configure their own copy of the kernel.
#define H2SVD_SET_SMEM(ctx, kern, bytes)                                                                   \
    do {                                                                                                   \
        static uint64_t _done_mask = 0; /* bit d: configured on device d (devices >= 64 are always reconfigured) */ \
        const int _d = (ctx)->device;                                                                      \
        if (_d >= 64 || !((_done_mask >> _d) & 1ull)) {                                                    \
            H2SVD_CUDA(cudaFuncSetAttribute((kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (_d < 64) _done_mask |= 1ull << _d;                                                         \
        }                                                                                                  \
    } while (0)

#define H2SVD_TRY(expr)                 \
    do {                                \
        int _rc = (expr);               \
        if (_rc != H2SVD_OK) return _rc; \
    } while (0)

using fr::Fr;

static inline const Fr* as_fr(const h2svd_fr* p) { return reinterpret_cast<const Fr*>(p); }
static inline Fr* as_fr(h2svd_fr* p) { return reinterpret_cast<Fr*>(p); }

// ---- device-side load/store of one field element as two 128-bit accesses ---------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ Fr ld_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// streaming store: witness arrays are written once and read back by the host, never re-read here
__device__ __forceinline__ void st_fr_cs(Fr* p, const Fr& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    __stcs(q, make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]));
    __stcs(q + 1, make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]));
}
#endif

// ---- kernel launchers (one per .cu) -----------------------------------------------------------------
int launch_fr_matmul(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m);
int launch_fr_matmul_naive(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k,
                           size_t m);
// tensor-core engine of the mat-mul (matmul_tc.cu)
bool fr_matmul_tc_supported(size_t n, size_t k, size_t m);
namespace rs { struct RescaleConsts; }
// fuse != nullptr: the epilogue also writes the rescale_matrix witnesses of C (out_q, out_wit as in launch_rescale)
int launch_fr_matmul_tc(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m,
                        const rs::RescaleConsts* fuse = nullptr, Fr* out_q = nullptr, Fr* out_wit = nullptr);
// mat-mul followed by rescale_matrix of the product: one fused launch when the tensor-core engine applies
int launch_fr_matmul_rescale(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m, int P, int lb,
                             int S, int A, Fr* out_q, Fr* out_wit);
int launch_transpose(h2svd_ctx* ctx, const Fr* src, Fr* dst, size_t rows, size_t cols);
int launch_gamma_powers(h2svd_ctx* ctx, const Fr* gamma, size_t d, Fr* out);
// totals (optional): last running sum of every row
int launch_mat_vec_prefix(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t len,
                          size_t v_row_stride, Fr* out, Fr* totals);
int launch_mat_vec_prefix2(h2svd_ctx* ctx, const Fr* a0, size_t rows0, Fr* out0, Fr* totals0, const Fr* a1,
                           size_t rows1, Fr* out1, Fr* totals1, const Fr* v, size_t len);
// row totals only, lazily accumulated (no running sums)
int launch_mat_vec_totals(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t len, Fr* totals);
int launch_gather(h2svd_ctx* ctx, const Fr* src, size_t count, size_t stride, size_t offset,
                  Fr* out);
int launch_is_equal(h2svd_ctx* ctx, const Fr* x, const Fr* y, size_t count, Fr* diff, Fr* is_zero,
                    Fr* inv);
int launch_rescale(h2svd_ctx* ctx, const Fr* cs, size_t count, int P, int lb, int S, int A, Fr* out_q,
                   Fr* out_wit);
int abs_less_than_params(const Fr& bnd, int lb, int with_diff, int* n_out, Fr* bound_out);
int launch_abs_less_than(h2svd_ctx* ctx, const Fr* x, const Fr* y, size_t count, const Fr& bnd, int lb, Fr* out_wit);
int range_check_params(int range_bits, int lb, int* n_out, int* rem_out);
int launch_range_check(h2svd_ctx* ctx, const Fr* x, size_t count, int range_bits, int lb, Fr* out_wit);
int launch_mat_times_diag(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t lda, size_t cols_v, Fr* out);
int launch_sub(h2svd_ctx* ctx, const Fr* a, const Fr* b, size_t count, Fr* out);
int launch_isqrt(h2svd_ctx* ctx, const Fr* a, size_t count, int P, Fr* out);
int launch_quantize(h2svd_ctx* ctx, const double* x, size_t count, int P, Fr* out);
int launch_check_canonical(h2svd_ctx* ctx, const Fr* x, size_t count, int* d_flag);
int launch_microbench(h2svd_ctx* ctx, int kind, int iters, double* ops_per_s);
int launch_microbench_hbm(h2svd_ctx* ctx, int kind, size_t bytes, double* gb_per_s);
int launch_microbench_i8(h2svd_ctx* ctx, int kind, double min_seconds, double* ops_per_s);

}  // namespace h2svd
