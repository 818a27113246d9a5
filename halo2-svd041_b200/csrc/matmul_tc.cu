// K1t: tensor-core engine of the Fr mat-mul  C = A * B  (reference src/matrix/mod.rs:507-535
// `honest_prover_mat_mul`; same contract as the IMAD kernels of matmul.cu, bit-exact with them).
//
// A field element in Montgomery form is 32 bytes.  Writing a = sum_p a_p 2^(8p), b = sum_q b_q 2^(8q),
//     sum_k a_ik * b_kj  =  sum_d 2^(8d) * D_d(i,j),      D_d(i,j) = sum_{p+q=d} sum_k a_p(i,k) * b_q(k,j),
// i.e. 63 "diagonal" sums of u8 x u8 products: exactly what the 5th-generation tensor cores compute
// (tcgen05.mma kind::i8, unsigned 8-bit operands, 32-bit accumulators in TMEM).  For k <= 1024 every D_d is below
// 32 * 1024 * 255^2 < 2^31, so the 32-bit accumulators are exact and the result is the same integer the IMAD
// kernels accumulate; one Montgomery reduction per C element (fr::reduce_wide_acc) finishes it.
//
// Mapping (one CTA per SM, persistent over tiles of 128 rows x 8 columns of C):
//   * MMA M = 128 rows i of A; for ONE byte plane p at a time the A operand is the 128 x K byte matrix a_p(i,k).
//   * MMA N = 256 = 32 byte planes q x 8 columns j of B, ordered q-major: row q*8+j of the B operand is b_q(k,j).
//   * the product for plane p is accumulated at TMEM column offset 8p: column (q*8+j) + 8p = (p+q)*8 + j = d*8 + j,
//     so the 32 planes p overlap-add straight into the 63 diagonals (x 8 columns j = 504 of the 512 TMEM columns).
//     Nothing but zero-initialised accumulators and `accumulate` MMAs is needed for the convolution structure.
//   * per 128 bytes of K: one TMA load of the B operand (32 KB, shared by all p) and 32 TMA loads of A planes
//     (16 KB each, 6-stage ring), 4 MMAs (K = 32 bytes each) per A plane: 128 x 256 x 32 MACs per instruction,
//     the full-rate shape (128 cycles per instruction per SM).
//   * epilogue (16 warps: thread = row i, 4 warps per TMEM lane quarter with 2 columns j each): reads the 63
//     diagonals of one (i, j) from TMEM, carries them into an 18-limb integer, reduces, stores C; then re-zeroes
//     its own accumulator columns for the next tile.
// Operands are pre-split into byte planes by two O(N^2) kernels (A8[p][i][k], B8[q][j][k], both K-major so that
// TMA with the 128-byte swizzle delivers the canonical K-major UMMA layout).
//
// Arithmetic per C element and k: 1024 u8 MACs on the tensor pipe (vs 48 IMAD.WIDE on the integer pipe).
#include <cuda.h>  // CUtensorMap and enums only; the encoder comes from cudaGetDriverEntryPoint (no -lcuda)

#include "common.cuh"
#include "rescale_dev.cuh"

namespace h2svd {

namespace {

constexpr int TC_BM = 128;       // rows of C per tile = MMA M = TMEM lanes
constexpr int TC_BJ = 8;         // columns of C per tile
constexpr int TC_BKB = 128;      // bytes (= k values) of K per pipeline unit: one 128-byte swizzle span
constexpr int TC_SA_PLAIN = 6;   // A-plane stages (6 x 16 KB + 2 x 32 KB leaves room for a co-resident mat-vec CTA)
#ifndef TC_SA_FUSED_CFG
#define TC_SA_FUSED_CFG 5
#endif
#ifndef TC_CH_CFG
#define TC_CH_CFG 4
#endif
constexpr int TC_SA_FUSED = TC_SA_FUSED_CFG;   // one stage less when the epilogue stages rescale witnesses in shared memory
constexpr int TC_SB = 2;         // B buffers
constexpr int TC_KB_PASS = 8;    // K blocks per accumulation pass: 1024 k values keep every diagonal below 2^31
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BKB;        // 16 KB
constexpr uint32_t TC_B_BYTES = 32 * TC_BJ * TC_BKB;   // 32 KB
// Epilogue warps: 4 per TMEM lane quarter with 2 of the 8 columns j each.  (Measured for the fused-rescale variant: 3 per
// quarter with 3/3/2 columns and 128 spill-free registers is slower, 1.03 vs 0.91 ms at N=1024 -- the code below still
// handles MAXC = 3.)
__host__ __device__ constexpr int tc_epi_warps(bool) { return 16; }
__host__ __device__ constexpr int tc_threads(bool fused) { return 64 + 32 * tc_epi_warps(fused); }  // + TMA producer warp + MMA issuer warp
// fused rescale: every epilogue warp stages 4 witnesses (128 B + 16 B skew) per lane, single-buffered
using TcWitnessStream = rs::WitnessStreamT<TC_CH_CFG, 1>;
constexpr uint32_t TC_STAGE_BYTES = tc_epi_warps(true) * 32 * TcWitnessStream::ROW_U4 * 16;
static_assert((size_t)TC_SA_FUSED_CFG * 16384 + 65536 + 1280 + TC_STAGE_BYTES <= 232448, "fused kernel: shared memory");
constexpr size_t tc_smem_bytes(bool fused) {
    return (size_t)(fused ? TC_SA_FUSED : TC_SA_PLAIN) * TC_A_BYTES + (size_t)TC_SB * TC_B_BYTES + 256 + 1024 +
           (fused ? TC_STAGE_BYTES : 0);
}
// kind::i8 instruction descriptor: D = S32, A = B = unsigned 8-bit, both K-major, N = 256, M = 128
constexpr uint32_t TC_IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (the launch fails loudly) instead of hanging the GPU.
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity, int* err) {
    if (tc_mbar_try(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!tc_mbar_try(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) {  // 10 s
            atomicExch(err, 3);
            __trap();
        }
    }
}
__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                          uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
// K-major, 128-byte-swizzled operand tile (rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024u >> 4) << 32;        // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld2(uint32_t taddr, uint32_t& v0, uint32_t& v1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(taddr));
}
// dg[0..63] (dg[63] = 0): the 63 diagonal sums of one C element  ->  T[0..17] = sum_d dg[d] * 2^(8d)
__device__ __forceinline__ void tc_carry_diagonals(const uint32_t* dg, uint32_t* T) {
    uint64_t cy = 0;
#pragma unroll
    for (int w = 0; w < 18; w++) {
        const uint32_t y0 = w < 16 ? dg[4 * w] : 0u;
        const uint32_t y1 = w < 16 ? dg[4 * w + 1] : 0u, y1p = (w >= 1 && w <= 16) ? dg[4 * w - 3] : 0u;
        const uint32_t y2 = w < 16 ? dg[4 * w + 2] : 0u, y2p = (w >= 1 && w <= 16) ? dg[4 * w - 2] : 0u;
        const uint32_t y3 = w < 16 ? dg[4 * w + 3] : 0u, y3p = (w >= 1 && w <= 16) ? dg[4 * w - 1] : 0u;
        cy += y0;
        cy += __funnelshift_l(y1p, y1, 8);
        cy += __funnelshift_l(y2p, y2, 16);
        cy += __funnelshift_l(y3p, y3, 24);
        T[w] = (uint32_t)cy;
        cy >>= 32;
    }
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t& v0, uint32_t& v1, uint32_t& v2, uint32_t& v3) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_st2_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(z), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_st1_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(z) : "memory");
}
// this warp's accumulators (its TMEM lanes, its cnt = 2 or 3 columns j of every diagonal) := 0, then hand them (back)
// to the MMA warp.  Each warp zeroes only its own columns, so the warps of a lane quarter never race.
__device__ __forceinline__ void tc_zero_and_release(uint32_t tcol0, int cnt, uint32_t bar) {
#pragma unroll 9
    for (int d = 0; d < 63; d++) {
        tc_st2_zero(tcol0 + 8u * d);
        if (cnt == 3) tc_st1_zero(tcol0 + 8u * d + 2);  // warp-uniform
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    tc_mbar_arrive(bar);
}

// ---- operand split: byte planes, K-major ---------------------------------------------------------------
// a: n x k (row-major Fr)  ->  planes[p][i][kk], kk < ldk (bytes kk >= k are zero); one thread = 4 k values
__global__ void tc_split_a_kernel(const Fr* __restrict__ a, uint32_t* __restrict__ planes, int n, int k, int ldk4) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * ldk4) return;
    const int i = (int)(idx / ldk4), w = (int)(idx % ldk4);
    Fr e[4];
#pragma unroll
    for (int t = 0; t < 4; t++) e[t] = (4 * w + t < k) ? ldg_fr(a + (size_t)i * k + 4 * w + t) : fr::zero();
#pragma unroll
    for (int p = 0; p < 32; p++) {
        const int limb = p >> 2, sh = (p & 3) * 8;
        const uint32_t word = ((e[0].l[limb] >> sh) & 0xffu) | (((e[1].l[limb] >> sh) & 0xffu) << 8) |
                              (((e[2].l[limb] >> sh) & 0xffu) << 16) | (((e[3].l[limb] >> sh) & 0xffu) << 24);
        planes[((size_t)p * n + i) * ldk4 + w] = word;
    }
}
// b: k x m (row-major Fr)  ->  planes[q][j][kk]  (the transposed operand: K-major rows per column j of B)
__global__ void tc_split_b_kernel(const Fr* __restrict__ b, uint32_t* __restrict__ planes, int k, int m, int ldk4) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)m * ldk4) return;
    const int j = (int)(idx / ldk4), w = (int)(idx % ldk4);
    Fr e[4];
#pragma unroll
    for (int t = 0; t < 4; t++) e[t] = (4 * w + t < k) ? ldg_fr(b + (size_t)(4 * w + t) * m + j) : fr::zero();
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int limb = q >> 2, sh = (q & 3) * 8;
        const uint32_t word = ((e[0].l[limb] >> sh) & 0xffu) | (((e[1].l[limb] >> sh) & 0xffu) << 8) |
                              (((e[2].l[limb] >> sh) & 0xffu) << 16) | (((e[3].l[limb] >> sh) & 0xffu) << 24);
        planes[((size_t)q * m + j) * ldk4 + w] = word;
    }
}

// ---- the tensor-core kernel ----------------------------------------------------------------------------
// FUSE (experimental, off by default -- h2svd_debug_set_fuse_rescale): the epilogue also emits the rescale_matrix
// witnesses (K4, rescale_dev.cuh) of every C element it produces, so that the witness stream (2 KB per element) is
// written under the MMAs of the next tile instead of after the mat-mul.  Bit-identical, but measured NOT to pay: the
// MMAs read their operands from shared memory at 96 of the SM's 128 B/clk, and staging the witnesses for the bulk
// stores needs another ~60 B/clk, so a fused tile takes 126 us against 69 us (MMAs) + 58 us (stand-alone rescale):
// N=1024 0.91 ms fused vs 1.01 ms as two kernels (a wave-rounding gain only), slower on 1-4 wave slabs.  Direct 32-byte
// global stores instead of staging: 1.37 ms.  It would need the A operand in TMEM or 2-CTA MMAs (halved B reads).
template <bool FUSE>
__global__ void __launch_bounds__(tc_threads(FUSE), 1)
fr_matmul_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    Fr* __restrict__ c, int n, int k, int m, int tiles_j, int num_tiles, int* err,
                    const __grid_constant__ rs::RescaleConsts kc, Fr* __restrict__ out_q, Fr* __restrict__ out_wit) {
    constexpr int TC_SA = FUSE ? TC_SA_FUSED : TC_SA_PLAIN;
    constexpr int EPI_WARPS = 16;
    constexpr int MAXC = TC_BJ / (EPI_WARPS / 4);  // columns j per epilogue warp (the last warp of a quarter may own fewer)
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc_smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // swizzle atoms need 1024-byte alignment
    uint8_t* smem = tc_smem_raw + (base - raw);
    const uint32_t s_a = base, s_b = base + TC_SA * TC_A_BYTES;
    const uint32_t bars = s_b + TC_SB * TC_B_BYTES;
    const uint32_t full_a = bars, empty_a = bars + 8 * TC_SA, full_b = bars + 16 * TC_SA,
                   empty_b = full_b + 8 * TC_SB, tmem_full = empty_b + 8 * TC_SB, tmem_empty = tmem_full + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (tmem_empty + 8 - base));
    uint4* stage = reinterpret_cast<uint4*>(smem + (bars + 256 - base));  // FUSE only: witness staging rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_SA; s++) {
            tc_mbar_init(full_a + 8 * s, 1);
            tc_mbar_init(empty_a + 8 * s, 1);
        }
        for (int s = 0; s < TC_SB; s++) {
            tc_mbar_init(full_b + 8 * s, 1);
            tc_mbar_init(empty_b + 8 * s, 1);
        }
        tc_mbar_init(tmem_full, 1);
        tc_mbar_init(tmem_empty, 32 * EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    const int kblocks = (k + TC_BKB - 1) / TC_BKB;
    const int passes = (kblocks + TC_KB_PASS - 1) / TC_KB_PASS;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t ua = 0, ub = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int ib = tile / tiles_j, jb = tile % tiles_j;
                for (int kb = 0; kb < kblocks; kb++) {
                    const uint32_t sb = ub % TC_SB;
                    tc_mbar_wait(empty_b + 8 * sb, ((ub / TC_SB) & 1) ^ 1, err);
                    tc_mbar_expect_tx(full_b + 8 * sb, TC_B_BYTES);
                    tc_tma_3d(s_b + sb * TC_B_BYTES, &tm_b, kb * TC_BKB, jb * TC_BJ, 0, full_b + 8 * sb);
                    ub++;
                    for (int p = 0; p < 32; p++) {
                        const uint32_t sa = ua % TC_SA;
                        tc_mbar_wait(empty_a + 8 * sa, ((ua / TC_SA) & 1) ^ 1, err);
                        tc_mbar_expect_tx(full_a + 8 * sa, TC_A_BYTES);
                        tc_tma_2d(s_a + sa * TC_A_BYTES, &tm_a, kb * TC_BKB, p * n + ib * TC_BM, full_a + 8 * sa);
                        ua++;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t ua = 0, ub = 0, round = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int pass = 0; pass < passes; pass++) {
                tc_mbar_wait(tmem_empty, round & 1, err);  // accumulators zeroed by the epilogue warps
                tc_fence_after();
                const int kb_end = min(kblocks, (pass + 1) * TC_KB_PASS);
                for (int kb = pass * TC_KB_PASS; kb < kb_end; kb++) {
                    const uint32_t sb = ub % TC_SB;
                    tc_mbar_wait(full_b + 8 * sb, (ub / TC_SB) & 1, err);
                    const uint64_t bdesc = tc_smem_desc(s_b + sb * TC_B_BYTES);
                    for (int p = 0; p < 32; p++) {
                        const uint32_t sa = ua % TC_SA;
                        tc_mbar_wait(full_a + 8 * sa, (ua / TC_SA) & 1, err);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t adesc = tc_smem_desc(s_a + sa * TC_A_BYTES);
#pragma unroll
                            for (int s = 0; s < TC_BKB / 32; s++)  // 32 bytes of K per instruction: +2 in 16-byte units
                                tc_mma_i8(tmem_base + 8u * p, adesc + 2u * s, bdesc + 2u * s, 1u);
                            tc_commit(empty_a + 8 * sa);  // frees the A stage once these MMAs have read it
                        }
                        __syncwarp();
                        ua++;
                    }
                    if (lane == 0) tc_commit(empty_b + 8 * sb);
                    __syncwarp();
                    ub++;
                }
                if (lane == 0) tc_commit(tmem_full);
                __syncwarp();
                round++;
            }
        }
    } else {
        // ===== epilogue: thread = row of the tile =====
        const uint32_t quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int il = quarter * 32 + lane;
        // this warp's columns j0 .. j0+cnt-1 of the tile: 2 each with 16 warps, 3/3/2 with 12 warps
        const int jg = (warp - 2) >> 2;
        const int j0 = jg * MAXC;
        const int cnt = TC_BJ - j0 < MAXC ? TC_BJ - j0 : MAXC;
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        tc_zero_and_release(tlane + j0, cnt, tmem_empty);
        TcWitnessStream ws;
        if (FUSE) {
            ws.warp_row0 = stage + (size_t)(warp - 2) * 32 * TcWitnessStream::ROW_U4;
            ws.row0 = ws.warp_row0 + (size_t)lane * TcWitnessStream::ROW_U4;
            ws.W = m * kc.p.W;
            ws.buf = 0;
            ws.fill = 0;
        }
        uint32_t round = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int ib = tile / tiles_j, jb = tile % tiles_j;
            const int gi = ib * TC_BM + il;
            for (int pass = 0; pass < passes; pass++) {
                tc_mbar_wait(tmem_full, round & 1, err);
                tc_fence_after();
                // Phase 1 (on the critical path of the next tile): read this warp's columns of all 63 diagonals and
                // carry them into 18-limb integers.  T = sum_d dg[d] * 2^(8d): diagonals d = 4g + r sit at whole-word
                // offsets g for fixed r, so T = Y0 + (Y1 << 8) + (Y2 << 16) + (Y3 << 24) with Y_r[g] = dg[4g + r].
                uint32_t dg[MAXC][64];
#pragma unroll
                for (int d = 0; d < 63; d++) {
                    if (MAXC == 2) {
                        tc_ld2(tlane + 8u * d + j0, dg[0][d], dg[1][d]);
                    } else {
                        uint32_t unused;  // 4-column load; a 2-column warp reads (and ignores) its neighbours' columns
                        tc_ld4(tlane + 8u * d + j0, dg[0][d], dg[1][d], dg[MAXC - 1][d], unused);
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // pin the loaded registers behind the wait (the compiler must not read them earlier)
#pragma unroll
                for (int d = 0; d < 63; d++)
#pragma unroll
                    for (int q = 0; q < MAXC; q++) asm volatile("" : "+r"(dg[q][d]));
                uint32_t T[MAXC][18];
#pragma unroll
                for (int q = 0; q < MAXC; q++) {
                    dg[q][63] = 0;
                    tc_carry_diagonals(dg[q], T[q]);
                }
                // the accumulators are free again: zero them and let the MMA warp start the next tile while the
                // Montgomery reductions (and the rescale witnesses) below run
                tc_fence_before();  // order the TMEM reads above before the zeroing stores / the next MMAs
                tc_zero_and_release(tlane + j0, cnt, tmem_empty);
                // Phase 2 (overlaps the next tile's MMAs)
                Fr res[MAXC];
#pragma unroll
                for (int q = 0; q < MAXC; q++) {
                    res[q] = fr::reduce_wide_acc(T[q]);
                    const int gj = jb * TC_BJ + j0 + q;
                    if (q < cnt && gi < n && gj < m) {
                        Fr* dst = c + (size_t)gi * m + gj;
                        if (pass > 0) res[q] = fr::add(ld_fr(dst), res[q]);
                        st_fr(dst, res[q]);
                    }
                }
                if (FUSE && pass == passes - 1) {
                    // this warp's 32 lanes hold rows row0 .. row0+31 of one column: stripes m*W witnesses apart
                    const int row0 = ib * TC_BM + (int)quarter * 32;
                    const int valid = n - row0 < 32 ? n - row0 : 32;
#pragma unroll
                    for (int q = 0; q < MAXC; q++) {
                        const int gj = jb * TC_BJ + j0 + q;
                        if (q < cnt && gj < m && valid > 0) {  // warp-uniform
                            ws.valid = valid;
                            ws.gwarp = out_wit + ((size_t)row0 * m + gj) * (size_t)kc.p.W;
                            const Fr qv = rs::rescale_element(ws, kc, gi < n ? res[q] : fr::zero());
                            if (gi < n) st_fr(out_q + (size_t)gi * m + gj, qv);
                        }
                    }
                }
                round++;
            }
        }
        // shared memory must outlive every bulk read, and the witness writes must be complete at kernel end
        if (FUSE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tc_encode_fn>(p);
    }
    return fn;
}

}  // namespace

int g_matmul_tc = -1;  // -1 auto, 0 off, 1 force (triage hook; see launch_fr_matmul)

bool fr_matmul_tc_supported(size_t n, size_t k, size_t m) {
    // p * n + ib * 128 and the plane sizes are 32-bit TMA coordinates / comfortably below 2^31
    return k >= 1 && n * 32 < (1ull << 31) && m * 32 < (1ull << 31) && k < (1ull << 30);
}

int launch_fr_matmul_tc(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m,
                        const rs::RescaleConsts* fuse, Fr* out_q, Fr* out_wit) {
    if (n == 0 || m == 0) return H2SVD_OK;
    tc_encode_fn encode = tc_encoder();
    if (!encode) {
        set_error("fr_matmul (tensor-core engine): cuTensorMapEncodeTiled is not available from this driver");
        return H2SVD_ECUDA;
    }
    const size_t ldk = (k + 15) & ~(size_t)15;  // TMA row pitch: multiple of 16 bytes
    const size_t bytes_a = 32 * n * ldk, bytes_b = 32 * m * ldk;
    const size_t need = ((bytes_a + 255) & ~(size_t)255) + bytes_b;
    if (ctx->kara_ws_bytes < need) {  // the operand workspace is shared with the Karatsuba engine (never both at once)
        H2SVD_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->kara_ws) H2SVD_CUDA(cudaFree(ctx->kara_ws));
        ctx->kara_ws = nullptr;
        ctx->kara_ws_bytes = 0;
        H2SVD_CUDA(cudaMalloc(&ctx->kara_ws, need));
        ctx->kara_ws_bytes = need;
    }
    uint8_t* a8 = reinterpret_cast<uint8_t*>(ctx->kara_ws);
    uint8_t* b8 = a8 + ((bytes_a + 255) & ~(size_t)255);
    const int ldk4 = (int)(ldk / 4);
    {
        const size_t ta = n * (size_t)ldk4, tb = m * (size_t)ldk4;
        tc_split_a_kernel<<<(unsigned)((ta + 255) / 256), 256, 0, ctx->stream>>>(a, reinterpret_cast<uint32_t*>(a8),
                                                                                 (int)n, (int)k, ldk4);
        H2SVD_LAUNCH_CHECK(ctx);
        tc_split_b_kernel<<<(unsigned)((tb + 255) / 256), 256, 0, ctx->stream>>>(b, reinterpret_cast<uint32_t*>(b8),
                                                                                 (int)k, (int)m, ldk4);
        H2SVD_LAUNCH_CHECK(ctx);
    }
    CUtensorMap tm_a, tm_b;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)(32 * n)};
        const cuuint64_t strides[1] = {(cuuint64_t)ldk};
        const cuuint32_t box[2] = {(cuuint32_t)TC_BKB, (cuuint32_t)TC_BM};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&tm_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, a8, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("fr_matmul (tensor-core engine): tensor map for A failed (CUresult %d)", (int)r);
            return H2SVD_ECUDA;
        }
    }
    {
        const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)m, 32};
        const cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)(m * ldk)};
        const cuuint32_t box[3] = {(cuuint32_t)TC_BKB, (cuuint32_t)TC_BJ, 32};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tm_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, b8, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("fr_matmul (tensor-core engine): tensor map for B failed (CUresult %d)", (int)r);
            return H2SVD_ECUDA;
        }
    }
    const int tiles_i = (int)((n + TC_BM - 1) / TC_BM), tiles_j = (int)((m + TC_BJ - 1) / TC_BJ);
    const long long tiles = (long long)tiles_i * tiles_j;
    if (tiles >= (1LL << 31)) {
        set_error("fr_matmul (tensor-core engine): too many tiles");
        return H2SVD_EINVAL;
    }
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    if (fuse) {
        H2SVD_SET_SMEM(ctx, fr_matmul_tc_kernel<true>, tc_smem_bytes(true));
        fr_matmul_tc_kernel<true><<<grid, tc_threads(true), tc_smem_bytes(true), ctx->stream>>>(
            tm_a, tm_b, c, (int)n, (int)k, (int)m, tiles_j, (int)tiles, ctx->d_flag, *fuse, out_q, out_wit);
    } else {
        static const rs::RescaleConsts none{};
        H2SVD_SET_SMEM(ctx, fr_matmul_tc_kernel<false>, tc_smem_bytes(false));
        fr_matmul_tc_kernel<false><<<grid, tc_threads(false), tc_smem_bytes(false), ctx->stream>>>(
            tm_a, tm_b, c, (int)n, (int)k, (int)m, tiles_j, (int)tiles, ctx->d_flag, none, nullptr, nullptr);
    }
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd

extern "C" int h2svd_debug_set_matmul_tc(int v) {
    h2svd::g_matmul_tc = v;
    return 0;
}
