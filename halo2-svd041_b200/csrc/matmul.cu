// K1: C = A * B over BN254-Fr  (replaces reference src/matrix/mod.rs:510-537 field_mat_mul): the integer-pipe (IMAD)
// engines and the dispatch between them and the tensor-core engine of matmul_tc.cu (launch_fr_matmul, at the end).
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
#include <algorithm>

#include "common.cuh"
#include "fr_acc.cuh"
#include "fr_kara.cuh"
#include "rescale_dev.cuh"

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

    const int warp = tid >> 5;
    for (int c = 0; c < nchunks; c++) {
        // rotating issuer: warp (c mod 8) stages chunk c + PD (no single warp carries the issue work of every chunk)
        if (warp == (c & (WARPS - 1)) && c + PD < nchunks) {
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

// ---------------------------------------------------------------------------------------------------
// Stream-K variant for shapes whose tile count does not fill the resident CTA slots evenly (e.g. the
// 128 x 1024 row slab of one of 8 GPUs at N=1024: 256 tiles on 296 slots = 86 % of a wave).
// The (tile, k-chunk) iteration space is flattened into `total_units` and cut into gridDim.x equal
// contiguous ranges, one per resident CTA, so every SM gets the same number of k-chunks (+-1).  The TMA
// pipeline runs straight through tile boundaries.  A CTA that covers a whole tile writes C directly; a
// partial tile goes, Montgomery-reduced, to this CTA's slot 0 (first segment) or slot 1 (last segment) of
// `partial`, and fr_matmul_fixup_kernel adds the 2-3 partials of every split tile.  No CTA ever waits on
// another one.  Exact: field addition of canonical partial sums is order-independent.
// CTA-wide state of the stream-K kernel, kept in shared memory so that the segment routine below has the
// same small register state as the one-CTA-per-tile kernel.
template <int STAGES>
struct SkShared {
    uint64_t full_bar[STAGES];
    uint64_t empty_bar[STAGES];
    // per-stage unit descriptor written by the issuing warp: {klen, row0, col0, kc} (consumers use klen)
    int4 meta[STAGES];
    const Fr* A;
    const Fr* B;
    Fr* C;
    Fr* partial;  // this CTA's two partial-tile slots
    int n, k, m, tiles_x, nchunks, u0, nloc, tile0;
};

// The one SkShared object of the CTA (static shared memory has a compile-time address, so neither it nor
// the dynamic-shared tile buffers cost a register in the segment routine).
template <int STAGES>
__device__ __forceinline__ SkShared<STAGES>* sk_shared() {
    __shared__ __align__(16) SkShared<STAGES> sh;
    return &sh;
}
extern __shared__ __align__(128) unsigned char sk_smem_raw[];

// warp 0 only: stage unit i (in-range rows / columns / k only) and publish its descriptor
template <int TM, int TN, int BK, int STAGES>
__device__ __forceinline__ void sk_issue_unit(int i, int lane) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    SkShared<STAGES>* sh = sk_shared<STAGES>();
    Fr* sA = reinterpret_cast<Fr*>(sk_smem_raw);
    Fr* sB = sA + STAGES * cfg::A_STAGE;
    const int n = sh->n, k = sh->k, m = sh->m, nchunks = sh->nchunks, tiles_x = sh->tiles_x, u0 = sh->u0;
    const int u = u0 + i;
    const int tile = u / nchunks, kc = u - tile * nchunks;
    const int trow = tile / tiles_x;
    const int row0 = trow * cfg::BM, col0 = (tile - trow * tiles_x) * cfg::BN;
    const int rows_valid = min(cfg::BM, n - row0), cols_valid = min(cfg::BN, m - col0);
    const int s = i % STAGES;
    const int k0 = kc * BK;
    const int klen = min(BK, k - k0);
    if (lane == 0) {
        sh->meta[s] = make_int4(klen, row0, col0, kc);
        const uint32_t bytes = (uint32_t)((rows_valid * klen + klen * cols_valid) * sizeof(Fr));
        mbar_arrive_expect_tx(&sh->full_bar[s], bytes);  // release: meta[s] is visible to whoever waits on it
    }
    __syncwarp();
    Fr* dA = sA + s * cfg::A_STAGE;
    Fr* dB = sB + s * cfg::B_STAGE;
    const Fr* A = sh->A;
    const Fr* B = sh->B;
    for (int r = lane; r < rows_valid; r += 32)
        tma_bulk_g2s(dA + r * BK, A + (size_t)(row0 + r) * k + k0, (uint32_t)(klen * sizeof(Fr)), &sh->full_bar[s]);
    for (int r = lane; r < klen; r += 32)
        tma_bulk_g2s(dB + r * cfg::BN, B + (size_t)(k0 + r) * m + col0, (uint32_t)(cols_valid * sizeof(Fr)),
                     &sh->full_bar[s]);
}

// One segment = this CTA's share of one tile, starting at local unit i; returns the next local unit.
// Deliberately NOT inlined, a COUNTED unit loop, and nothing but the accumulators live across the k loop
// (the tile coordinates are recomputed after it): with any other shape ptxas stops keeping the 72 accumulator
// registers in aligned pairs and shuffles them with IMAD.MOV / XOR swaps every k step -- up to +70 %
// instructions on the very pipe the kernel is bound by (measured 98 vs 129 G mul-add/s).
template <int TM, int TN, int BK, int STAGES>
__device__ __noinline__ int sk_segment(int i) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    constexpr int PD = STAGES - 2;
    SkShared<STAGES>* sh = sk_shared<STAGES>();
    Fr* sA = reinterpret_cast<Fr*>(sk_smem_raw);
    Fr* sB = sA + STAGES * cfg::A_STAGE;
    fr::WideAcc acc[TM][TN];
#pragma unroll
    for (int ii = 0; ii < TM; ii++)
#pragma unroll
        for (int j = 0; j < TN; j++) fr::acc_clear(acc[ii][j]);
    // units of this segment: to the end of the tile or of this CTA's range, whichever comes first
    const int kc0 = (sh->u0 + i) % sh->nchunks;
    const int cnt = min(sh->nchunks - kc0, sh->nloc - i);
    int s = 0;
    const Fr* pA0 = sA + (threadIdx.x / TX) * BK;   // hoisted: fewer live temporaries in the unit loop
    const Fr* pB0 = sB + (threadIdx.x % TX);
    for (int c = 0; c < cnt; c++, i++) {
        // rotating issuer (see sk_segment_k): warp (i mod 8) stages unit i + PD
        if ((int)(threadIdx.x >> 5) == (i & (WARPS - 1)) && i + PD < sh->nloc) {
            if (i >= 2) mbar_wait(&sh->empty_bar[(i - 2) % STAGES], ((i - 2) / STAGES) & 1);
            sk_issue_unit<TM, TN, BK, STAGES>(i + PD, (int)(threadIdx.x & 31));
        }
        s = i % STAGES;
        mbar_wait(&sh->full_bar[s], (i / STAGES) & 1);
        const Fr* pA = pA0 + s * cfg::A_STAGE;
        const Fr* pB = pB0 + s * cfg::B_STAGE;
        const int klen = sh->meta[s].x;
#pragma unroll 1
        for (int kk = 0; kk < klen; kk++) {
            Fr a[TM], b[TN];
#pragma unroll
            for (int ii = 0; ii < TM; ii++) a[ii] = ld_fr(pA + ii * TY * BK + kk);
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = ld_fr(pB + kk * cfg::BN + j * TX);
#pragma unroll
            for (int ii = 0; ii < TM; ii++)
#pragma unroll
                for (int j = 0; j < TN; j++) fr::mul_acc(acc[ii][j], a[ii].l, b[j].l);
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&sh->empty_bar[s]);
    }
    // where this segment lives (recomputed here rather than carried through the k loop in registers)
    const int u_last = sh->u0 + i - 1;
    const int tile = u_last / sh->nchunks, kc_last = u_last - tile * sh->nchunks;
    const int trow = tile / sh->tiles_x;
    const int row0 = trow * cfg::BM, col0 = (tile - trow * sh->tiles_x) * cfg::BN;
    const int flags = ((kc_last == sh->nchunks - 1 && cnt == sh->nchunks) ? 2 : 0) | ((tile == sh->tile0 ? 0 : 1) << 2);

    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const bool whole = (flags & 2) != 0;
    const int n = sh->n, m = sh->m;
    Fr* C = sh->C;
    Fr* pdst = sh->partial + (size_t)(flags >> 2) * (cfg::BM * cfg::BN);
#pragma unroll
    for (int ii = 0; ii < TM; ii++) {
        const int rl = ty + ii * TY;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int cl = tx + j * TX;
            if (row0 + rl < n && col0 + cl < m) {
                const Fr val = fr::acc_finalize(acc[ii][j]);
                if (whole)
                    st_fr(C + (size_t)(row0 + rl) * m + col0 + cl, val);
                else
                    st_fr(pdst + rl * cfg::BN + cl, val);
            }
        }
    }
    return i;
}

template <int TM, int TN, int BK, int STAGES, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
fr_matmul_streamk_kernel(const Fr* __restrict__ A, const Fr* __restrict__ B, Fr* __restrict__ C,
                         Fr* __restrict__ partial, int n, int k, int m, int tiles_x, int nchunks,
                         long long total_units) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    constexpr int PD = STAGES - 2;
    SkShared<STAGES>* sh = sk_shared<STAGES>();
    const int tid = threadIdx.x;

    if (tid == 0) {
        // total_units < 2^31 (checked by the launcher): per-unit index math stays 32-bit
        const int u0 = (int)(((long long)blockIdx.x * total_units) / gridDim.x);
        const int u1 = (int)(((long long)(blockIdx.x + 1) * total_units) / gridDim.x);
        sh->A = A; sh->B = B; sh->C = C;
        sh->partial = partial + (size_t)blockIdx.x * 2 * (cfg::BM * cfg::BN);
        sh->n = n; sh->k = k; sh->m = m; sh->tiles_x = tiles_x; sh->nchunks = nchunks;
        sh->u0 = u0; sh->nloc = u1 - u0; sh->tile0 = u0 / nchunks;
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sh->full_bar[s], 1);
            mbar_init(&sh->empty_bar[s], WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nloc = sh->nloc;
    if (tid < 32) {
        for (int i = 0; i < PD && i < nloc; i++) sk_issue_unit<TM, TN, BK, STAGES>(i, tid);
    }
    int i = 0;
    while (i < nloc) i = sk_segment<TM, TN, BK, STAGES>(i);
}

// Adds the partial sums of every tile that stream-K split over several CTAs (see above).
template <int BM, int BN>
__global__ void fr_matmul_fixup_kernel(Fr* __restrict__ C, const Fr* __restrict__ partial, int n, int m, int tiles_x,
                                       int nchunks, long long total_units, int G) {
    const int tile = blockIdx.x;
    const long long ub = (long long)tile * nchunks, ue = ub + nchunks - 1;
    auto owner = [&](long long u) { return (int)(((u + 1) * G + total_units - 1) / total_units - 1); };
    const int c_first = owner(ub), c_last = owner(ue);
    if (c_first == c_last) return;  // written directly by its only CTA
    const int row0 = (tile / tiles_x) * BM, col0 = (tile % tiles_x) * BN;
    for (int e = threadIdx.x; e < BM * BN; e += blockDim.x) {
        const int rl = e / BN, cl = e % BN;
        if (row0 + rl >= n || col0 + cl >= m) continue;
        Fr sum = fr::zero();
        for (int c = c_first; c <= c_last; c++) {
            const long long u0 = ((long long)c * total_units) / G;
            const int slot = (u0 / nchunks == tile) ? 0 : 1;  // this tile is the CTA's first segment, else its last
            sum = fr::add(sum, ldg_fr(partial + ((size_t)c * 2 + slot) * (BM * BN) + e));
        }
        st_fr(C + (size_t)(row0 + rl) * m + col0 + cl, sum);
    }
}

// ---------------------------------------------------------------------------------------------------
// Karatsuba variant (fr_kara.cuh): 48 instead of 64 IMAD.WIDE per multiply-add on pre-split operands.
// Same TMA/mbarrier pipeline as fr_matmul_kernel; 16x16 C tile, ONE C element per thread (three 4x4-limb lazy
// accumulators = 57 registers), operands in the 48-byte {lo, hi, s} layout produced by kara_split_kernel.
using fr::KOp;

__global__ void kara_split_kernel(const Fr* __restrict__ src, KOp* __restrict__ dst, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const KOp k = fr::ksplit(ldg_fr(src + i));
        uint4* q = reinterpret_cast<uint4*>(dst + i);
        q[0] = make_uint4(k.lo[0], k.lo[1], k.lo[2], k.lo[3]);
        q[1] = make_uint4(k.hi[0], k.hi[1], k.hi[2], k.hi[3]);
        q[2] = make_uint4(k.s[0], k.s[1], k.s[2], k.s[3]);
    }
}

template <int BK, int STAGES>
struct KCfg {
    static constexpr int BM = TY, BN = TX;
    static constexpr int A_STAGE = BM * BK;  // KOp elements
    static constexpr int B_STAGE = BK * BN;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(KOp);
};

template <int BK, int STAGES, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
fr_matmul_kara_kernel(const KOp* __restrict__ A, const KOp* __restrict__ B, Fr* __restrict__ C, int n, int k, int m) {
    using cfg = KCfg<BK, STAGES>;
    static_assert(STAGES >= 3, "need >= 3 stages");
    constexpr int PD = STAGES - 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    KOp* sA = reinterpret_cast<KOp*>(smem_raw);
    KOp* sB = sA + STAGES * cfg::A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool is_issuer = tid < 32;
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

    auto issue_chunk = [&](int c) {
        const int s = c % STAGES;
        const int k0 = c * BK;
        const int klen = min(BK, k - k0);
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((rows_valid * klen + klen * cols_valid) * sizeof(KOp));
            mbar_arrive_expect_tx(&full_bar[s], bytes);
        }
        __syncwarp();
        KOp* dA = sA + s * cfg::A_STAGE;
        KOp* dB = sB + s * cfg::B_STAGE;
        for (int r = lane; r < rows_valid; r += 32)
            tma_bulk_g2s(dA + r * BK, A + (size_t)(row0 + r) * k + k0, (uint32_t)(klen * sizeof(KOp)), &full_bar[s]);
        for (int r = lane; r < klen; r += 32)
            tma_bulk_g2s(dB + r * cfg::BN, B + (size_t)(k0 + r) * m + col0, (uint32_t)(cols_valid * sizeof(KOp)),
                         &full_bar[s]);
    };

    if (is_issuer) {
        for (int c = 0; c < PD && c < nchunks; c++) issue_chunk(c);
    }

    const int tx = tid % TX, ty = tid / TX;
    fr::KAcc p0, p1, p2;
    fr::kacc_clear(p0);
    fr::kacc_clear(p1);
    fr::kacc_clear(p2);

    const int warp = tid >> 5;
    for (int c = 0; c < nchunks; c++) {
        // rotating issuer: warp (c mod 8) stages chunk c + PD, so no single warp carries the TMA issue work of every chunk
        if (warp == (c & (WARPS - 1)) && c + PD < nchunks) {
            if (c >= 2) mbar_wait(&empty_bar[(c - 2) % STAGES], ((c - 2) / STAGES) & 1);
            issue_chunk(c + PD);
        }
        const int s = c % STAGES;
        mbar_wait(&full_bar[s], (c / STAGES) & 1);
        const uint4* pA = reinterpret_cast<const uint4*>(sA + s * cfg::A_STAGE + ty * BK);  // + 3*kk
        const uint4* pB = reinterpret_cast<const uint4*>(sB + s * cfg::B_STAGE + tx);       // + 3*kk*BN
        const int klen = min(BK, k - c * BK);
#pragma unroll 1
        for (int kk = 0; kk < klen; kk++) {
            const uint4* qa = pA + 3 * kk;
            const uint4* qb = pB + 3 * kk * cfg::BN;
            {
                const uint4 a = qa[0], b = qb[0];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p0, av, bv);
            }
            {
                const uint4 a = qa[1], b = qb[1];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p2, av, bv);
            }
            {
                const uint4 a = qa[2], b = qb[2];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p1, av, bv);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    const int r = row0 + ty, cc = col0 + tx;
    if (r < n && cc < m) st_fr(C + (size_t)r * m + cc, fr::kara_finalize(p0, p1, p2));
}

// ---- the same stream-K schedule for the Karatsuba engine (KOp operands, three KAcc accumulators, 16x16 tiles) ----
template <int STAGES>
struct SkSharedK {
    uint64_t full_bar[STAGES];
    uint64_t empty_bar[STAGES];
    // per-stage unit descriptor written by the issuing warp: {klen, row0, col0, kc} (consumers use klen)
    int4 meta[STAGES];
    const KOp* A;
    const KOp* B;
    Fr* C;
    Fr* partial;  // this CTA's two partial-tile slots
    int n, k, m, tiles_x, nchunks, u0, nloc, tile0;
};

// The one SkShared object of the CTA (static shared memory has a compile-time address, so neither it nor
// the dynamic-shared tile buffers cost a register in the segment routine).
template <int STAGES>
__device__ __forceinline__ SkSharedK<STAGES>* sk_shared_k() {
    __shared__ __align__(16) SkSharedK<STAGES> sh;
    return &sh;
}

// warp 0 only: stage unit i (in-range rows / columns / k only) and publish its descriptor
template <int BK, int STAGES>
__device__ __forceinline__ void sk_issue_unit_k(int i, int lane) {
    using cfg = KCfg<BK, STAGES>;
    SkSharedK<STAGES>* sh = sk_shared_k<STAGES>();
    KOp* sA = reinterpret_cast<KOp*>(sk_smem_raw);
    KOp* sB = sA + STAGES * cfg::A_STAGE;
    const int n = sh->n, k = sh->k, m = sh->m, nchunks = sh->nchunks, tiles_x = sh->tiles_x, u0 = sh->u0;
    const int u = u0 + i;
    const int tile = u / nchunks, kc = u - tile * nchunks;
    const int trow = tile / tiles_x;
    const int row0 = trow * cfg::BM, col0 = (tile - trow * tiles_x) * cfg::BN;
    const int rows_valid = min(cfg::BM, n - row0), cols_valid = min(cfg::BN, m - col0);
    const int s = i % STAGES;
    const int k0 = kc * BK;
    const int klen = min(BK, k - k0);
    if (lane == 0) {
        sh->meta[s] = make_int4(klen, row0, col0, kc);
        const uint32_t bytes = (uint32_t)((rows_valid * klen + klen * cols_valid) * sizeof(KOp));
        mbar_arrive_expect_tx(&sh->full_bar[s], bytes);  // release: meta[s] is visible to whoever waits on it
    }
    __syncwarp();
    KOp* dA = sA + s * cfg::A_STAGE;
    KOp* dB = sB + s * cfg::B_STAGE;
    const KOp* A = sh->A;
    const KOp* B = sh->B;
    for (int r = lane; r < rows_valid; r += 32)
        tma_bulk_g2s(dA + r * BK, A + (size_t)(row0 + r) * k + k0, (uint32_t)(klen * sizeof(KOp)), &sh->full_bar[s]);
    for (int r = lane; r < klen; r += 32)
        tma_bulk_g2s(dB + r * cfg::BN, B + (size_t)(k0 + r) * m + col0, (uint32_t)(cols_valid * sizeof(KOp)),
                     &sh->full_bar[s]);
}

// One segment = this CTA's share of one tile, starting at local unit i; returns the next local unit.
// Deliberately NOT inlined, a COUNTED unit loop, and nothing but the accumulators live across the k loop
// (the tile coordinates are recomputed after it): with any other shape ptxas stops keeping the 72 accumulator
// registers in aligned pairs and shuffles them with IMAD.MOV / XOR swaps every k step -- up to +70 %
// instructions on the very pipe the kernel is bound by (measured 98 vs 129 G mul-add/s).
template <int BK, int STAGES>
__device__ __noinline__ int sk_segment_k(int i) {
    using cfg = KCfg<BK, STAGES>;
    constexpr int PD = STAGES - 2;
    SkSharedK<STAGES>* sh = sk_shared_k<STAGES>();
    KOp* sA = reinterpret_cast<KOp*>(sk_smem_raw);
    KOp* sB = sA + STAGES * cfg::A_STAGE;
    fr::KAcc p0, p1, p2;
    fr::kacc_clear(p0);
    fr::kacc_clear(p1);
    fr::kacc_clear(p2);
    // units of this segment: to the end of the tile or of this CTA's range, whichever comes first
    const int kc0 = (sh->u0 + i) % sh->nchunks;
    const int cnt = min(sh->nchunks - kc0, sh->nloc - i);
    int s = 0;
    const KOp* pA0 = sA + (threadIdx.x / TX) * BK;   // hoisted: fewer live temporaries in the unit loop
    const KOp* pB0 = sB + (threadIdx.x % TX);
    for (int c = 0; c < cnt; c++, i++) {
        // ROTATING issuer: warp (i mod 8) stages unit i + PD, so the ~150 instructions of coordinate arithmetic and TMA
        // issue are spread over all warps instead of making warp 0 the critical path of every unit (+12 % on one warp)
        if ((int)(threadIdx.x >> 5) == (i & (WARPS - 1)) && i + PD < sh->nloc) {
            if (i >= 2) mbar_wait(&sh->empty_bar[(i - 2) % STAGES], ((i - 2) / STAGES) & 1);
            sk_issue_unit_k<BK, STAGES>(i + PD, (int)(threadIdx.x & 31));
        }
        s = i % STAGES;
        mbar_wait(&sh->full_bar[s], (i / STAGES) & 1);
        const uint4* pA = reinterpret_cast<const uint4*>(pA0 + s * cfg::A_STAGE);
        const uint4* pB = reinterpret_cast<const uint4*>(pB0 + s * cfg::B_STAGE);
        const int klen = sh->meta[s].x;
#pragma unroll 1
        for (int kk = 0; kk < klen; kk++) {
            const uint4* qa = pA + 3 * kk;
            const uint4* qb = pB + 3 * kk * cfg::BN;
            {
                const uint4 a = qa[0], b = qb[0];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p0, av, bv);
            }
            {
                const uint4 a = qa[1], b = qb[1];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p2, av, bv);
            }
            {
                const uint4 a = qa[2], b = qb[2];
                const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                fr::kmul_acc(p1, av, bv);
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&sh->empty_bar[s]);
    }
    // where this segment lives (recomputed here rather than carried through the k loop in registers)
    const int u_last = sh->u0 + i - 1;
    const int tile = u_last / sh->nchunks, kc_last = u_last - tile * sh->nchunks;
    const int trow = tile / sh->tiles_x;
    const int row0 = trow * cfg::BM, col0 = (tile - trow * sh->tiles_x) * cfg::BN;
    const int flags = ((kc_last == sh->nchunks - 1 && cnt == sh->nchunks) ? 2 : 0) | ((tile == sh->tile0 ? 0 : 1) << 2);

    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const bool whole = (flags & 2) != 0;
    const int n = sh->n, m = sh->m;
    Fr* C = sh->C;
    Fr* pdst = sh->partial + (size_t)(flags >> 2) * (cfg::BM * cfg::BN);
    if (row0 + ty < n && col0 + tx < m) {
        const Fr val = fr::kara_finalize(p0, p1, p2);
        if (whole)
            st_fr(C + (size_t)(row0 + ty) * m + col0 + tx, val);
        else
            st_fr(pdst + ty * cfg::BN + tx, val);
    }
    return i;
}

template <int BK, int STAGES, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
fr_matmul_streamk_kara_kernel(const KOp* __restrict__ A, const KOp* __restrict__ B, Fr* __restrict__ C,
                         Fr* __restrict__ partial, int n, int k, int m, int tiles_x, int nchunks,
                         long long total_units) {
    using cfg = KCfg<BK, STAGES>;
    constexpr int PD = STAGES - 2;
    SkSharedK<STAGES>* sh = sk_shared_k<STAGES>();
    const int tid = threadIdx.x;

    if (tid == 0) {
        // total_units < 2^31 (checked by the launcher): per-unit index math stays 32-bit
        const int u0 = (int)(((long long)blockIdx.x * total_units) / gridDim.x);
        const int u1 = (int)(((long long)(blockIdx.x + 1) * total_units) / gridDim.x);
        sh->A = A; sh->B = B; sh->C = C;
        sh->partial = partial + (size_t)blockIdx.x * 2 * (cfg::BM * cfg::BN);
        sh->n = n; sh->k = k; sh->m = m; sh->tiles_x = tiles_x; sh->nchunks = nchunks;
        sh->u0 = u0; sh->nloc = u1 - u0; sh->tile0 = u0 / nchunks;
#pragma unroll
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sh->full_bar[s], 1);
            mbar_init(&sh->empty_bar[s], WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nloc = sh->nloc;
    if (tid < 32) {
        for (int i = 0; i < PD && i < nloc; i++) sk_issue_unit_k<BK, STAGES>(i, tid);
    }
    int i = 0;
    while (i < nloc) i = sk_segment_k<BK, STAGES>(i);
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

template <int TM, int TN, int BK, int STAGES, int MINBLOCKS>
int launch_variant(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, int n, int k, int m) {
    using cfg = Cfg<TM, TN, BK, STAGES>;
    const int tiles_x = (m + cfg::BN - 1) / cfg::BN, tiles_y = (n + cfg::BM - 1) / cfg::BM;
    const long long tiles = (long long)tiles_x * tiles_y;
    const int nchunks = (k + BK - 1) / BK;
    const int slots = ctx->sm_count * MINBLOCKS;
    // Schedule choice (measured on B200, tools/matmul_bench.py): the one-CTA-per-tile launch runs at
    // 129 G mul-add/s once there are >= ~3.5 waves of tiles (a partial last wave is softened because an SM
    // holding one CTA instead of two runs it ~1.6x faster), but drops to 106-110 G/s at 0.9-1.7 waves (the
    // 128- and 256-row slabs of the 8- and 4-GPU split of N=1024); stream-K holds ~118 G/s regardless.
    const bool streamk = ctx->tune.streamk < 0 ? (tiles * 2 < 5LL * slots && nchunks >= 4 && tiles * nchunks >= 2LL * slots)
                                       : ctx->tune.streamk != 0;
    if (streamk && nchunks >= 1 && tiles * nchunks < (1LL << 31)) {
        auto kern = fr_matmul_streamk_kernel<TM, TN, BK, STAGES, MINBLOCKS>;
        H2SVD_SET_SMEM(ctx, kern, cfg::SMEM);
        const long long total_units = tiles * nchunks;
        const int G = (int)(total_units < slots ? total_units : slots);
        const size_t part_bytes = (size_t)G * 2 * cfg::BM * cfg::BN * sizeof(Fr);
        H2SVD_TRY(ws_grow(ctx, &ctx->sk_ws, &ctx->sk_ws_bytes, part_bytes));
        kern<<<G, THREADS, cfg::SMEM, ctx->stream>>>(a, b, c, (Fr*)ctx->sk_ws, n, k, m, tiles_x, nchunks, total_units);
        H2SVD_LAUNCH_CHECK(ctx);
        fr_matmul_fixup_kernel<cfg::BM, cfg::BN><<<(unsigned)tiles, 256, 0, ctx->stream>>>(
            c, (const Fr*)ctx->sk_ws, n, m, tiles_x, nchunks, total_units, G);
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    auto kern = fr_matmul_kernel<TM, TN, BK, STAGES, MINBLOCKS>;
    H2SVD_SET_SMEM(ctx, kern, cfg::SMEM);
    dim3 grid(tiles_x, tiles_y);
    kern<<<grid, THREADS, cfg::SMEM, ctx->stream>>>(a, b, c, n, k, m);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace

template <int BK, int STAGES, int MINBLOCKS>
static int launch_kara(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, int n, int k, int m) {
    using cfg = KCfg<BK, STAGES>;
    auto kern = fr_matmul_kara_kernel<BK, STAGES, MINBLOCKS>;
    H2SVD_SET_SMEM(ctx, kern, cfg::SMEM);
    // pre-split operands (48 bytes per element) in a dedicated workspace
    const size_t na = (size_t)n * k, nb = (size_t)k * m;
    const size_t bytes = (na + nb) * sizeof(KOp);
    H2SVD_TRY(ws_grow(ctx, &ctx->kara_ws, &ctx->kara_ws_bytes, bytes));
    KOp* ka = (KOp*)ctx->kara_ws;
    KOp* kb = ka + na;
    const unsigned sb = (unsigned)std::min<size_t>((na + 255) / 256, (size_t)ctx->sm_count * 8);
    kara_split_kernel<<<sb, 256, 0, ctx->stream>>>(a, ka, na);
    H2SVD_LAUNCH_CHECK(ctx);
    const unsigned sb2 = (unsigned)std::min<size_t>((nb + 255) / 256, (size_t)ctx->sm_count * 8);
    kara_split_kernel<<<sb2, 256, 0, ctx->stream>>>(b, kb, nb);
    H2SVD_LAUNCH_CHECK(ctx);
    const int tiles_x = (m + cfg::BN - 1) / cfg::BN, tiles_y = (n + cfg::BM - 1) / cfg::BM;
    const long long tiles = (long long)tiles_x * tiles_y;
    const int nchunks = (k + BK - 1) / BK;
    const int slots = ctx->sm_count * MINBLOCKS;
    const bool streamk = ctx->tune.streamk < 0 ? (tiles * 2 < 7LL * slots && nchunks >= 4 && tiles * nchunks >= 2LL * slots)
                                       : ctx->tune.streamk != 0;
    if (streamk && tiles * nchunks < (1LL << 31)) {
        auto skern = fr_matmul_streamk_kara_kernel<BK, STAGES, MINBLOCKS>;
        H2SVD_SET_SMEM(ctx, skern, cfg::SMEM);
        const long long total_units = tiles * nchunks;
        const int G = (int)(total_units < slots ? total_units : slots);
        const size_t part_bytes = (size_t)G * 2 * cfg::BM * cfg::BN * sizeof(Fr);
        H2SVD_TRY(ws_grow(ctx, &ctx->sk_ws, &ctx->sk_ws_bytes, part_bytes));
        skern<<<G, THREADS, cfg::SMEM, ctx->stream>>>(ka, kb, c, (Fr*)ctx->sk_ws, n, k, m, tiles_x, nchunks, total_units);
        H2SVD_LAUNCH_CHECK(ctx);
        fr_matmul_fixup_kernel<cfg::BM, cfg::BN><<<(unsigned)tiles, 256, 0, ctx->stream>>>(
            c, (const Fr*)ctx->sk_ws, n, m, tiles_x, nchunks, total_units, G);
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    dim3 grid(tiles_x, tiles_y);
    kern<<<grid, THREADS, cfg::SMEM, ctx->stream>>>(ka, kb, c, n, k, m);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

// Tensor-core engine (matmul_tc.cu: u8 byte planes on tcgen05, ~12x the IMAD engines at N=1024; measured with
// tools/tc_check.py: ahead from 64^3 on, behind for very short k where its per-element epilogue dominates).
// Forcing one of the IMAD engines / schedules through the triage hooks switches the automatic choice off.
static bool use_tensor_engine(const h2svd_ctx* ctx, size_t n, size_t k, size_t m) {
    const int g_matmul_tc = ctx->tune.matmul_tc;
    const bool imad_forced = ctx->tune.kara >= 0 || ctx->tune.streamk >= 0 || ctx->tune.variant != 0;
    // a handful of 128 x 8 tiles with a long k (e.g. 64 x 4096 x 64: 0.31 vs 0.15 ms) is the one skinny shape the IMAD
    // engines win: their stream-K schedule splits k over the SMs, the tensor-core kernel walks it tile by tile
    const size_t tiles = ((n + 127) / 128) * ((m + 7) / 8);
    const bool few_long = tiles * 8 < (size_t)ctx->sm_count && k > 2048;
    const bool tc = g_matmul_tc == 1 ||
                    (g_matmul_tc < 0 && !imad_forced && !few_long && k >= 32 && n * k * m >= (1ull << 18));
    return tc && fr_matmul_tc_supported(n, k, m);
}

int launch_fr_matmul_rescale(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m, int P, int lb,
                             int S, int A, Fr* out_q, Fr* out_wit) {
    if (n == 0 || m == 0) return H2SVD_OK;
    rs::RescaleConsts kc;
    const int W = make_rescale_consts(P, lb, S, A, &kc);
    if (ctx->tune.fuse_rescale > 0 && W > 0 && m * (size_t)W < (1ull << 31) && n <= (1u << 30) && m <= (1u << 30) &&
        k <= (1u << 30) && use_tensor_engine(ctx, n, k, m)) {
        ctx->last_engine = 2;
        return launch_fr_matmul_tc(ctx, a, b, c, n, k, m, &kc, out_q, out_wit);
    }
    H2SVD_TRY(launch_fr_matmul(ctx, a, b, c, n, k, m));
    return launch_rescale(ctx, c, n * m, P, lb, S, A, out_q, out_wit);
}

int launch_fr_matmul(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m) {
    if (n == 0 || m == 0) return H2SVD_OK;
    if (n > (1u << 30) || m > (1u << 30) || k > (1u << 30)) {
        set_error("fr_matmul: dimension too large");
        return H2SVD_EINVAL;
    }
    if (use_tensor_engine(ctx, n, k, m)) {
        ctx->last_engine = 2;
        return launch_fr_matmul_tc(ctx, a, b, c, n, k, m);
    }
    const int g_kara = ctx->tune.kara, g_variant = ctx->tune.variant;
    ctx->last_engine = (g_kara >= 1 || (g_kara < 0 && k >= 64 && n * m >= 4096)) ? 1 : 0;
    // Karatsuba engine (48 instead of 64 IMAD.WIDE per multiply-add; measured 147 vs 129 G mul-add/s at N=1024) unless
    // the product is too small for the O(N^2) operand split and the two extra launches to pay off
    if (g_kara == 1) return launch_kara<16, 3, 2>(ctx, a, b, c, (int)n, (int)k, (int)m);
    if (g_kara == 2) return launch_kara<16, 3, 3>(ctx, a, b, c, (int)n, (int)k, (int)m);
    if (g_kara == 3 || (g_kara < 0 && k >= 64 && n * m >= 4096))
        return launch_kara<16, 4, 2>(ctx, a, b, c, (int)n, (int)k, (int)m);
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

