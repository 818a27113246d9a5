// K1: C = A * B over BN254-Fr  (replaces reference src/matrix/mod.rs:510-537 field_mat_mul).
//
// Design (sm_100a, integer-pipe bound -- see DESIGN.md "K1"):
//  * CTA tile (16*TM) x (16*TN) of C, 256 consumer threads in a 16x16 layout, each owning a TM x TN
//    register tile of lazy 18-limb accumulators (fr_acc.cuh): 64 IMAD.WIDE.U32 + 16 IADD3.X per Fr
//    multiply-add, ZERO Montgomery reductions inside the k loop; one reduction per C element at the end.
//  * A and B k-slabs are staged in shared memory by the TMA engine (cp.async.bulk, one bulk copy per
//    tile row, completion on an mbarrier), STAGES deep; no thread ever issues a global load and there
//    is no __syncthreads in the main loop.
//  * Warps release a stage with one mbarrier arrive each; warp 0 re-fills it two chunks later.
//  * Ragged shapes: the producer copies only in-range rows/columns, the consumers loop only over
//    in-range k and skip out-of-range stores; nothing is ever read or written out of bounds.
#include "common.cuh"
#include "fr_acc.cuh"

namespace h2svd {

namespace {

constexpr int TY = 16, TX = 16;  // thread layout inside a CTA
constexpr int THREADS = TY * TX;
constexpr int WARPS = THREADS / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// TMA 1-D bulk copy global -> shared, completion signalled on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int TM, int TN, int BK, int STAGES>
struct Cfg {
    static constexpr int BM = TY * TM, BN = TX * TN;
    static constexpr int A_STAGE = BM * BK;  // Fr elements
    static constexpr int B_STAGE = BK * BN;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(Fr);
};

// Pipeline: STAGES smem buffers, prefetch distance STAGES-2.  Warp 0 doubles as the TMA issuer: at
// the top of k-chunk c it refills the buffer that chunk c-2 used (every warp released that one at
// least a whole chunk of work ago, so the wait practically never blocks) with chunk c+STAGES-2.
template <int TM, int TN, int BK, int STAGES, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
fr_matmul_kernel(const Fr* __restrict__ A, const Fr* __restrict__ B, Fr* __restrict__ C, int n, int k, int m) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    static_assert(STAGES >= 3, "need >= 3 stages");
    constexpr int PD = STAGES - 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    Fr* sA = reinterpret_cast<Fr*>(smem_raw);
    Fr* sB = sA + STAGES * cfg::A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool is_issuer = tid < 32;  // warp 0
    const int row0 = blockIdx.y * cfg::BM;
    const int col0 = blockIdx.x * cfg::BN;
    const int nchunks = (k + BK - 1) / BK;
    const int rows_valid = min(cfg::BM, n - row0);
    const int cols_valid = min(cfg::BN, m - col0);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0 only: stage k-chunk c into buffer c % STAGES (in-range rows / columns / k only)
    auto issue_chunk = [&](int c) {
        const int s = c % STAGES;
        const int k0 = c * BK;
        const int klen = min(BK, k - k0);
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((rows_valid * klen + klen * cols_valid) * sizeof(Fr));
            mbar_arrive_expect_tx(&full_bar[s], bytes);
        }
        __syncwarp();
        Fr* dA = sA + s * cfg::A_STAGE;
        Fr* dB = sB + s * cfg::B_STAGE;
        for (int r = lane; r < rows_valid; r += 32)
            tma_bulk_g2s(dA + r * BK, A + (size_t)(row0 + r) * k + k0, (uint32_t)(klen * sizeof(Fr)),
                         &full_bar[s]);
        for (int r = lane; r < klen; r += 32)
            tma_bulk_g2s(dB + r * cfg::BN, B + (size_t)(k0 + r) * m + col0, (uint32_t)(cols_valid * sizeof(Fr)),
                         &full_bar[s]);
    };

    if (is_issuer) {
        for (int c = 0; c < PD && c < nchunks; c++) issue_chunk(c);
    }

    const int tx = tid % TX, ty = tid / TX;
    fr::WideAcc acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) fr::acc_clear(acc[i][j]);

    for (int c = 0; c < nchunks; c++) {
        if (is_issuer && c + PD < nchunks) {
            if (c >= 2) mbar_wait(&empty_bar[(c - 2) % STAGES], ((c - 2) / STAGES) & 1);
            issue_chunk(c + PD);
        }
        const int s = c % STAGES;
        mbar_wait(&full_bar[s], (c / STAGES) & 1);
        const Fr* pA = sA + s * cfg::A_STAGE + ty * BK;  // + i*TY*BK + kk
        const Fr* pB = sB + s * cfg::B_STAGE + tx;       // + kk*BN + j*TX
        const int klen = min(BK, k - c * BK);
#pragma unroll 1
        for (int kk = 0; kk < klen; kk++) {
            Fr a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = ld_fr(pA + i * TY * BK + kk);
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = ld_fr(pB + kk * cfg::BN + j * TX);
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) fr::mul_acc(acc[i][j], a[i].l, b[j].l);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ================= epilogue: one Montgomery reduction per C element =================
#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int r = row0 + ty + i * TY;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int cc = col0 + tx + j * TX;
            if (r < n && cc < m) st_fr(C + (size_t)r * m + cc, fr::acc_finalize(acc[i][j]));
        }
    }
}

// One thread per C element, fully reduced arithmetic.  Debug/triage only (not on any product path).
__global__ void fr_matmul_naive_kernel(const Fr* __restrict__ A, const Fr* __restrict__ B, Fr* __restrict__ C,
                                       int n, int k, int m) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * m) return;
    const int i = (int)(idx / m), j = (int)(idx % m);
    Fr acc = fr::zero();
    for (int t = 0; t < k; t++) acc = fr::add(acc, fr::mont_mul(ldg_fr(A + (size_t)i * k + t), ldg_fr(B + (size_t)t * m + j)));
    st_fr(C + idx, acc);
}

__global__ void transpose_kernel(const Fr* __restrict__ src, Fr* __restrict__ dst, int rows, int cols) {
    // 32x32 tile through shared memory, both sides coalesced
    __shared__ uint4 tile[2][32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int y = by + r, x = bx + threadIdx.x;
        if (y < rows && x < cols) {
            const uint4* p = reinterpret_cast<const uint4*>(src + (size_t)y * cols + x);
            tile[0][r][threadIdx.x] = p[0];
            tile[1][r][threadIdx.x] = p[1];
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int y = bx + r, x = by + threadIdx.x;  // dst is cols x rows
        if (y < cols && x < rows) {
            uint4* p = reinterpret_cast<uint4*>(dst + (size_t)y * rows + x);
            p[0] = tile[0][threadIdx.x][r];
            p[1] = tile[1][threadIdx.x][r];
        }
    }
}

int g_variant = 0;

template <int TM, int TN, int BK, int STAGES, int MINBLOCKS>
int launch_variant(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, int n, int k, int m) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    auto kern = fr_matmul_kernel<TM, TN, BK, STAGES, MINBLOCKS>;
    static bool configured = false;  // per process; attribute is per function, device-wide
    if (!configured) {
        H2SVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg::SMEM));
        configured = true;
    }
    dim3 grid((m + cfg::BN - 1) / cfg::BN, (n + cfg::BM - 1) / cfg::BM);
    kern<<<grid, THREADS, cfg::SMEM, ctx->stream>>>(a, b, c, n, k, m);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace

int launch_fr_matmul(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m) {
    if (n == 0 || m == 0) return H2SVD_OK;
    if (n > (1u << 30) || m > (1u << 30) || k > (1u << 30)) {
        set_error("fr_matmul: dimension too large");
        return H2SVD_EINVAL;
    }
    switch (g_variant) {
        case 1: return launch_variant<1, 2, 16, 3, 2>(ctx, a, b, c, (int)n, (int)k, (int)m);
        case 2: return launch_variant<1, 1, 16, 3, 3>(ctx, a, b, c, (int)n, (int)k, (int)m);
        case 3: return launch_variant<2, 2, 16, 3, 1>(ctx, a, b, c, (int)n, (int)k, (int)m);
        default: return launch_variant<2, 1, 16, 3, 2>(ctx, a, b, c, (int)n, (int)k, (int)m);
    }
}

int launch_fr_matmul_naive(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m) {
    if (n == 0 || m == 0) return H2SVD_OK;
    const size_t total = n * m;
    fr_matmul_naive_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(a, b, c, (int)n, (int)k, (int)m);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_transpose(h2svd_ctx* ctx, const Fr* src, Fr* dst, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return H2SVD_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(src, dst, (int)rows, (int)cols);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd

extern "C" int h2svd_debug_set_matmul_variant(int v) {
    h2svd::g_variant = v;
    return 0;
}
