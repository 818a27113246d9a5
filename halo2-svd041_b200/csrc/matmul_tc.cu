// K1t: tensor-core engines of the Fr mat-mul  C = A * B  (reference src/matrix/mod.rs:507-535
// `honest_prover_mat_mul`; same contract as the IMAD kernels of matmul.cu, bit-exact with them).
//
// FULL-WIDTH engine (TcFull).  A field element in Montgomery form is 32 bytes.  Writing a = sum_p a_p 2^(8p),
// b = sum_q b_q 2^(8q),
//     sum_k a_ik * b_kj  =  sum_d 2^(8d) * D_d(i,j),      D_d(i,j) = sum_{p+q=d} sum_k a_p(i,k) * b_q(k,j),
// i.e. 63 "diagonal" sums of u8 x u8 products: exactly what the 5th-generation tensor cores compute
// (tcgen05.mma kind::i8, unsigned 8-bit operands, 32-bit accumulators in TMEM).  For k <= 1024 every D_d is below
// 32 * 1024 * 255^2 < 2^31, so the 32-bit accumulators are exact and the result is the same integer the IMAD
// kernels accumulate; one Montgomery reduction per C element (fr::reduce_wide_acc) finishes it.
//
// SMALL-OPERAND engine (TcSmall, tc_small.cuh).  Quantized fixed-point operands (ZkMatrix::new, :230-252) are small SIGNED
// integers in standard form (|q| < 2^70 covers P = 63 and |x| < 128).  The split kernels detect this on the device; the
// product is then taken over 9 x 9 balanced signed byte digits (s8 x s8; MMA N padded to a multiple of 16 by zero rows or a
// zero 10th B plane) -- 81 instead of 1024 byte products per multiply-add -- and Montgomery-encoded once per C element.  Same bytes
// out.  If any operand is out of range the full-width engine runs instead: both engines are enqueued, a device flag
// written by the split kernels decides which one does the work (no host round trip, so the sequence is graph-capturable).
//
// Mapping (one CTA per SM, persistent over tiles of 128 rows x BJ columns of C; BJ = 8 full-width, 8-28 small):
//   * MMA M = 128 rows i of A; for ONE byte plane p at a time the A operand is the 128 x K byte matrix a_p(i,k).
//   * MMA N = LB byte planes q x BJ columns j of B, ordered q-major: row q*BJ+j of the B operand is b_q(k,j).
//   * the product for plane p is accumulated at TMEM column offset BJ*p: column (q*BJ+j) + BJ*p = (p+q)*BJ + j = d*BJ + j,
//     so the planes p overlap-add straight into the diagonals (63 x 8 = 504 / 18 x 24 = 432 of the 512 TMEM columns).
//     Nothing but zero-initialised accumulators and `accumulate` MMAs is needed for the convolution structure.
//   * per 128 bytes of K: one TMA load of the B operand (shared by all p) and one TMA load per A plane
//     (16 KB each, multi-stage ring), 4 MMAs (K = 32 bytes each) per A plane.
//   * epilogue (16 warps: thread = row i, 4 warps per TMEM lane quarter with BJ/4 columns j each): reads the
//     diagonals of one (i, j) from TMEM, carries them into a multi-limb integer, reduces / encodes, stores C; the
//     accumulator columns are re-zeroed and handed back before the field arithmetic, which overlaps the next tile.
// Operands are pre-split into byte planes by O(N^2) kernels (A8[p][i][k], B8[q][j][k], both K-major so that
// TMA with the 128-byte swizzle delivers the canonical K-major UMMA layout).

#include "common.cuh"
#include "rescale_dev.cuh"
#include "tc_small.cuh"
#include "tma_util.cuh"

namespace h2svd {

namespace {

constexpr int TC_BM = 128;       // rows of C per tile = MMA M = TMEM lanes
constexpr int TC_BKB = 128;      // bytes (= k values) of K per pipeline unit: one 128-byte swizzle span
#ifndef TC_SA_FUSED_CFG
#define TC_SA_FUSED_CFG 5
#endif
#ifndef TC_CH_CFG
#define TC_CH_CFG 4
#endif
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BKB;        // 16 KB per A-plane stage
constexpr int TC_EPI_WARPS = 16;                        // 4 per TMEM lane quarter
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;      // + TMA producer warp + MMA issuer warp

// Full-width operands: 32 unsigned byte planes each; 1024 k values per accumulation pass keep every diagonal below 2^31.
struct TcFull {
    static constexpr int LA = 32, LB = 32, BJ = 8, SA = 6, SB = 2, KB_PASS = 8;   // 6 x 16 KB + 2 x 32 KB of stages
    static constexpr bool SIGNED = false;
};
// Small signed operands: 9 (A) x 9 or 10 (B) balanced byte digits (a zero 10th plane where MMA N = LB * BJ would not be a
// multiple of 16 otherwise); |D_d| <= 9 * k * 128^2 stays below 2^31 for k <= 14563: 64 K blocks (8192 k values) per pass.
// Four tile widths: 128 x 28 (MMA N = 252 + 4 zero rows: the fewest operand bytes per multiply-add), 128 x 24 (N = 240),
// 128 x 16 and 128 x 8 (N = 144 / 80: more tiles for the row slabs of a sharded job); tc_small_tile_width() picks.
template <int BJ_, int LB_, int SA_>
struct TcSmallT {
    static constexpr int LA = fr::SMALL_DIGITS, LB = LB_, BJ = BJ_, SA = SA_, SB = 2, KB_PASS = 64;
    static constexpr bool SIGNED = true;
};
// What bounds a tile (measured: tools/tc_timeline.py, tools/cluster_bench.py, h2svd_microbench_tensor_i8):
//   * MMAs alone (operands resident): N = 256 runs at 4.33-4.40 P op/s (128 clocks per instruction), N = 144 at 3.93 P
//     (81 clocks), N = 80 at 2.69 P (68 clocks -- the 128 x 32-byte A operand cannot be read faster).
//   * In the real kernel a tile's MMA phase takes 2.5-3.3 us per 1024 bytes of K for EVERY width (20-27 us at k = 1024):
//     each K block streams 9 x 16 KB of A planes + the B stage from L2 into each of the 148 SMs, ~32 B/clk per SM =
//     ~4700 B/clk for the chip, three quarters of the L2 throughput cap (~6300 B/clk).  7 and 10 A stages alike; clusters
//     of 2 or 4 CTAs multicasting each A stage alike (the L2 does not deduplicate multicast below 8 CTAs).
// Hence: the widest tile (most columns per A byte) wherever it does not cost a wave -- 28 columns use 252 of the MMA's
// 256 N (the 24-column tile 216 of 240: its 10th B plane is padding).
using TcSmall28 = TcSmallT<28, fr::SMALL_DIGITS, 7>;       // 7 x 16 KB + 2 x 32 KB: MMA N = 252 (+ 4 zero rows = 256)
using TcSmall = TcSmallT<24, fr::SMALL_DIGITS + 1, 7>;     // 7 x 16 KB + 2 x 30 KB of stages
using TcSmall16 = TcSmallT<16, fr::SMALL_DIGITS, 8>;       // 8 x 16 KB + 2 x 18 KB
using TcSmall8 = TcSmallT<8, fr::SMALL_DIGITS + 1, 8>;     // 8 x 16 KB + 2 x 10 KB
template <class C>
struct TcD {
    static constexpr int NROWS = C::LB * C::BJ;                 // rows of the B operand TMA delivers: (plane q, column j)
    static constexpr int NMMA = (NROWS + 15) / 16 * 16;         // MMA N; rows NROWS .. NMMA-1 of a B stage stay zero
    static constexpr int NDIAG = C::LA + C::LB - 1;             // diagonals d = p + q
    static constexpr int TCOLS = C::BJ * (C::LA - 1) + NMMA;    // TMEM columns the MMAs write (the padding rows add zeros)
    static constexpr int CW = ((C::BJ + 3) / 4 + 1) / 2 * 2;    // columns j per epilogue warp (even; the last group may hold fewer)
    static constexpr uint32_t B_BYTES = (uint32_t)NMMA * TC_BKB;   // stage size
    static constexpr uint32_t B_TX = (uint32_t)NROWS * TC_BKB;     // bytes one TMA load delivers
    // kind::i8 instruction descriptor: D = S32, A/B = unsigned (0) or signed (1) 8-bit, both K-major, N, M = 128
    static constexpr uint32_t IDESC = (2u << 4) | ((C::SIGNED ? 1u : 0u) << 7) | ((C::SIGNED ? 1u : 0u) << 10) |
                                      ((uint32_t)(NMMA >> 3) << 17) | ((128u >> 4) << 24);
    static_assert(NMMA % 16 == 0 && NMMA >= 16 && NMMA <= 256, "MMA N for M = 128");
    static_assert(TCOLS <= 512 && C::BJ % 4 == 0 && CW % 2 == 0 && 3 * CW < C::BJ && C::BJ <= 4 * CW, "TMEM columns / epilogue mapping");
    static_assert(B_BYTES % 1024 == 0, "B stages must keep the 1024-byte swizzle-atom alignment");
    static_assert((size_t)C::SA * TC_A_BYTES + (size_t)C::SB * B_BYTES + 256 + 1024 <= 232448, "stages exceed the SM's shared memory");
    static_assert(16 * C::SA + 16 * C::SB + 24 <= 256, "barrier block");
};
// fused rescale (8-column tiles of either engine): every epilogue warp stages 4 witnesses (128 B + 16 B skew) per lane, single-buffered
using TcWitnessStream = rs::WitnessStreamT<TC_CH_CFG, 1>;
constexpr uint32_t TC_STAGE_BYTES = TC_EPI_WARPS * 32 * TcWitnessStream::ROW_U4 * 16;
static_assert((size_t)TC_SA_FUSED_CFG * 16384 + 65536 + 1280 + TC_STAGE_BYTES <= 232448, "fused kernel: shared memory");
template <class C, bool FUSE>
__host__ __device__ constexpr int tc_sa() { return FUSE ? TC_SA_FUSED_CFG : C::SA; }
template <class C, bool FUSE>
constexpr size_t tc_smem_bytes() {
    return (size_t)tc_sa<C, FUSE>() * TC_A_BYTES + (size_t)C::SB * TcD<C>::B_BYTES + 256 + 1024 + (FUSE ? TC_STAGE_BYTES : 0);
}

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
// Cluster forms (CS > 1: the CTAs of a cluster work on the same 128 rows of A and different columns of B): one TMA load
// delivers its slice of an A-plane stage into the SAME shared-memory offset of every CTA of the cluster and signals the
// same barrier offset in each; one tcgen05.commit arrives on the same barrier offset in every CTA.
__device__ __forceinline__ void tc_tma_2d_multicast(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar,
                                                    uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tc_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
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
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
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
__device__ __forceinline__ void tc_st2_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(z), "r"(z) : "memory");
}
// debug timeline (triage switch "matmul_timeline"): CTA 0 stamps %globaltimer at the phase boundaries of its first 16 rounds
__device__ __forceinline__ void tc_stamp(unsigned long long* tl, uint32_t round, int slot) {
    if (tl != nullptr && blockIdx.x == 0 && round < 16) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        tl[8 * round + slot] = t;
    }
}

// this warp's accumulators (its TMEM lanes, its CW columns j of every diagonal) := 0, then hand them (back)
// to the MMA warp.  Each warp zeroes only its own columns, so the warps of a lane quarter never race.
template <class C>
__device__ __forceinline__ void tc_zero_and_release(uint32_t tcol0, uint32_t bar, int cw) {
#pragma unroll 9
    for (int d = 0; d < TcD<C>::NDIAG; d++) {
#pragma unroll
        for (int pr = 0; pr < TcD<C>::CW / 2; pr++)
            if (4 * TcD<C>::CW == C::BJ || 2 * pr < cw) tc_st2_zero(tcol0 + (uint32_t)(C::BJ * d + 2 * pr));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    tc_mbar_arrive(bar);
}

// ---- operand split: byte planes, K-major ---------------------------------------------------------------
// `mode` (device int, may be null) arbitrates between the two engines without a host round trip: the small-operand
// split kernels (SMALL) always run and raise *mode when an element is out of the small range; the full-width split
// kernels and the two mat-mul kernels read it and return at once when the other engine is the one to run.
// Four consecutive k values of one row / column -> one 32-bit word of every byte plane.
template <class C>
__device__ __forceinline__ void tc_plane_words(const Fr* e, uint32_t* words, int nplanes, bool* out_of_range) {
    if constexpr (C::SIGNED) {
        uint32_t t[4][3];
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 4; r++) ok &= fr::small_biased_fast(e[r], t[r]);   // zero (K padding) is in range: digits 0
        *out_of_range = !ok;
#pragma unroll
        for (int p = 0; p < 32; p++) {
            if (p < nplanes) {
                words[p] = p < fr::SMALL_DIGITS
                               ? (fr::small_digit_bits(t[0], p) | (fr::small_digit_bits(t[1], p) << 8) |
                                  (fr::small_digit_bits(t[2], p) << 16) | (fr::small_digit_bits(t[3], p) << 24))
                               : 0u;   // the padding plane of B
            }
        }
    } else {
        *out_of_range = false;
#pragma unroll
        for (int p = 0; p < 32; p++) {
            const int limb = p >> 2, sh = (p & 3) * 8;
            words[p] = ((e[0].l[limb] >> sh) & 0xffu) | (((e[1].l[limb] >> sh) & 0xffu) << 8) |
                       (((e[2].l[limb] >> sh) & 0xffu) << 16) | (((e[3].l[limb] >> sh) & 0xffu) << 24);
        }
    }
}

// ONE launch splits both operands: the first blocks_a CTAs take A, the others B.
//   a: n x k (row-major Fr)  ->  planes_a[p][i][kk], kk < ldk (bytes kk >= k are zero); one thread = 4 k values of a row.
//   b: k x m (row-major Fr)  ->  planes_b[q][j][kk]  (the transposed operand: K-major rows per column j of B).  One CTA =
//      128 k values x 8 columns: the loads walk B along its rows (8 x 32 B = 256 contiguous bytes per k), the words go
//      through a padded shared-memory tile, and every plane row leaves as one coalesced 128-byte store.
constexpr int TC_SPLIT_ROW_W = 36;   // words per (plane, column) tile row: 32 + 4 of skew (conflict-free both ways)
template <class C>
__global__ void __launch_bounds__(256)
tc_split_kernel(const Fr* __restrict__ a, uint32_t* __restrict__ planes_a, int n, const Fr* __restrict__ b,
                uint32_t* __restrict__ planes_b, int k, int m, int ldk4, int* mode, unsigned blocks_a, unsigned tiles_k) {
    if (!C::SIGNED && mode && *mode == 0) return;   // the small-operand engine runs: nothing to prepare
    __shared__ uint32_t tile[C::LB * 8 * TC_SPLIT_ROW_W];
    uint32_t words[32];
    bool bad;
    if (blockIdx.x < blocks_a) {
        const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (idx >= (size_t)n * ldk4) return;
        const int i = (int)(idx / ldk4), w = (int)(idx % ldk4);
        Fr e[4];
#pragma unroll
        for (int t = 0; t < 4; t++) e[t] = (4 * w + t < k) ? ldg_fr(a + (size_t)i * k + 4 * w + t) : fr::zero();
        tc_plane_words<C>(e, words, C::LA, &bad);
        if (C::SIGNED && bad) atomicOr(mode, 1);
#pragma unroll
        for (int p = 0; p < C::LA; p++) planes_a[((size_t)p * n + i) * ldk4 + w] = words[p];
        return;
    }
    const unsigned bb = blockIdx.x - blocks_a;
    const int k0 = (int)(bb % tiles_k) * 128, j0 = (int)(bb / tiles_k) * 8;
    const int j = threadIdx.x & 7, kq = threadIdx.x >> 3;       // kq: 0..31 -> k values k0 + 4*kq .. +3
    Fr e[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int kk = k0 + 4 * kq + t;
        e[t] = (kk < k && j0 + j < m) ? ldg_fr(b + (size_t)kk * m + j0 + j) : fr::zero();
    }
    tc_plane_words<C>(e, words, C::LB, &bad);
    if (C::SIGNED && bad) atomicOr(mode, 1);
#pragma unroll
    for (int q = 0; q < C::LB; q++) tile[(q * 8 + j) * TC_SPLIT_ROW_W + kq] = words[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w0 = k0 / 4;
    for (int row = warp; row < C::LB * 8; row += 8) {
        const int q = row >> 3, jj = row & 7;
        if (j0 + jj < m && w0 + lane < ldk4)
            planes_b[((size_t)q * m + j0 + jj) * ldk4 + w0 + lane] = tile[row * TC_SPLIT_ROW_W + lane];
    }
}

// ---- the tensor-core kernel ----------------------------------------------------------------------------
// FUSE (experimental, off by default -- tuning switch "fuse_rescale", full-width engine only): the epilogue also emits
// the rescale_matrix witnesses (K4, rescale_dev.cuh) of every C element it produces, so that the witness stream (2 KB per
// element) is written under the MMAs of the next tile instead of after the mat-mul.  Bit-identical, but measured NOT to
// pay: the MMAs read their operands from shared memory at 96 of the SM's 128 B/clk, and staging the witnesses for the
// bulk stores needs another ~60 B/clk, so a fused tile takes 126 us against 69 us (MMAs) + 58 us (stand-alone rescale).
template <class C, bool FUSE, int CS>
__global__ void __launch_bounds__(TC_THREADS, 1)
fr_matmul_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    Fr* __restrict__ c, int n, int k, int m, int tiles_j, int num_tiles, int* err,
                    const int* __restrict__ mode, int run_if_mode, unsigned long long* __restrict__ tl,
                    const __grid_constant__ rs::RescaleConsts kc, Fr* __restrict__ out_q, Fr* __restrict__ out_wit) {
    using D = TcD<C>;
    constexpr int TC_SA = tc_sa<C, FUSE>();
    constexpr int TC_SB = C::SB;
    constexpr int EPI_WARPS = TC_EPI_WARPS;
    constexpr int CW = D::CW;
    constexpr int BJ = C::BJ;
    constexpr uint32_t TC_B_BYTES = D::B_BYTES;
    // engine arbitration (uniform over the grid): the split kernels earlier in the stream decided which engine runs
    if (mode != nullptr && (*mode != 0) != (run_if_mode != 0)) return;
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
            tc_mbar_init(empty_a + 8 * s, CS);   // an A stage is refilled by every CTA of the cluster: all of them must have read it
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
    if constexpr (D::NMMA != D::NROWS) {
        // MMA N is rounded up to a multiple of 16: the rows of every B stage that TMA never writes are zero for the
        // whole kernel (their products land on the first columns of later diagonals, as zeros)
        constexpr int PAD_U4 = (D::NMMA - D::NROWS) * TC_BKB / 16;
        for (int i = threadIdx.x; i < TC_SB * PAD_U4; i += TC_THREADS)
            reinterpret_cast<uint4*>(smem + (s_b - base) + (size_t)(i / PAD_U4) * TC_B_BYTES + (size_t)D::NROWS * TC_BKB)[i % PAD_U4] =
                make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    // Cluster of CS CTAs: the same row block ib, CS consecutive column tiles.  crank = this CTA's column tile within the
    // group and its slice of every A-plane stage.  No CTA may multicast before every barrier of the cluster exists.
    const uint32_t crank = CS > 1 ? tc_cluster_rank() : 0u;
    constexpr uint16_t CMASK = (uint16_t)((1u << CS) - 1u);
    const int groups_j = (tiles_j + CS - 1) / CS;                  // column-tile groups per row block
    const int num_groups = (num_tiles / tiles_j) * groups_j;       // num_tiles = tiles_i * tiles_j
    const int cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
    if (CS > 1) tc_cluster_sync();

    const int kblocks = (k + TC_BKB - 1) / TC_BKB;
    const int passes = (kblocks + C::KB_PASS - 1) / C::KB_PASS;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t ua = 0, ub = 0, ptile = 0;
            for (int g = cluster_id; g < num_groups; g += num_clusters, ptile++) {
                // jb >= tiles_j (a group's surplus CTA): still loads its slice of A for the others; its own columns
                // are out of range (TMA zero-fills, the epilogue stores nothing)
                const int ib = g / groups_j, jb = (g % groups_j) * CS + (int)crank;
                tc_stamp(tl, ptile * passes, 6);
                for (int kb = 0; kb < kblocks; kb++) {
                    const uint32_t sb = ub % TC_SB;
                    tc_mbar_wait(empty_b + 8 * sb, ((ub / TC_SB) & 1) ^ 1, err);
                    tc_mbar_expect_tx(full_b + 8 * sb, D::B_TX);
                    tc_tma_3d(s_b + sb * TC_B_BYTES, &tm_b, kb * TC_BKB, jb * BJ, 0, full_b + 8 * sb);
                    ub++;
                    for (int p = 0; p < C::LA; p++) {
                        const uint32_t sa = ua % TC_SA;
                        tc_mbar_wait(empty_a + 8 * sa, ((ua / TC_SA) & 1) ^ 1, err);
                        tc_mbar_expect_tx(full_a + 8 * sa, TC_A_BYTES);
                        if constexpr (CS == 1) {
                            tc_tma_2d(s_a + sa * TC_A_BYTES, &tm_a, kb * TC_BKB, p * n + ib * TC_BM, full_a + 8 * sa);
                        } else {
                            // rows [crank * 128/CS, +128/CS) of the stage, into every CTA of the cluster (the tensor map's
                            // box is 128/CS rows; whole 1024-byte swizzle atoms)
                            constexpr int SL = TC_BM / CS;
                            tc_tma_2d_multicast(s_a + sa * TC_A_BYTES + crank * (SL * TC_BKB), &tm_a, kb * TC_BKB,
                                                p * n + ib * TC_BM + (int)crank * SL, full_a + 8 * sa, CMASK);
                        }
                        ua++;
                    }
                }
                tc_stamp(tl, ptile * passes, 7);
            }
            if constexpr (CS > 1) {
                // the other CTAs' MMA warps still signal this CTA's "empty" barriers: drain them before leaving
                for (int i = 0; i < TC_SA; i++, ua++)
                    tc_mbar_wait(empty_a + 8 * (ua % TC_SA), ((ua / TC_SA) & 1) ^ 1, err);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t ua = 0, ub = 0, round = 0;
        for (int g = cluster_id; g < num_groups; g += num_clusters) {
            for (int pass = 0; pass < passes; pass++) {
                tc_mbar_wait(tmem_empty, round & 1, err);  // accumulators zeroed by the epilogue warps
                tc_fence_after();
                if (lane == 0) tc_stamp(tl, round, 0);
                const int kb_end = min(kblocks, (pass + 1) * C::KB_PASS);
                for (int kb = pass * C::KB_PASS; kb < kb_end; kb++) {
                    const uint32_t sb = ub % TC_SB;
                    tc_mbar_wait(full_b + 8 * sb, (ub / TC_SB) & 1, err);
                    const uint64_t bdesc = tc_smem_desc(s_b + sb * TC_B_BYTES);
                    for (int p = 0; p < C::LA; p++) {
                        const uint32_t sa = ua % TC_SA;
                        tc_mbar_wait(full_a + 8 * sa, (ua / TC_SA) & 1, err);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t adesc = tc_smem_desc(s_a + sa * TC_A_BYTES);
#pragma unroll
                            for (int s = 0; s < TC_BKB / 32; s++)  // 32 bytes of K per instruction: +2 in 16-byte units
                                tc_mma_i8(tmem_base + (uint32_t)(BJ * p), adesc + 2u * s, bdesc + 2u * s, D::IDESC, 1u);
                            // frees the A stage once these MMAs have read it (in every CTA that refills it)
                            if constexpr (CS == 1) tc_commit(empty_a + 8 * sa);
                            else tc_commit_multicast(empty_a + 8 * sa, CMASK);
                        }
                        __syncwarp();
                        ua++;
                    }
                    if (lane == 0) tc_commit(empty_b + 8 * sb);
                    __syncwarp();
                    ub++;
                }
                if (lane == 0) {
                    tc_commit(tmem_full);
                    tc_stamp(tl, round, 1);
                }
                __syncwarp();
                round++;
            }
        }
    } else {
        // ===== epilogue: thread = row of the tile =====
        const uint32_t quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int il = quarter * 32 + lane;
        const int jg = (warp - 2) >> 2;     // this warp's columns j0 .. j0+cw-1 of the tile
        const int j0 = jg * CW;
        const int cw = BJ - j0 < CW ? BJ - j0 : CW;   // warp-uniform; < CW only in the last group of a 28-column tile
        const uint32_t tlane = tmem_base + ((quarter * 32u) << 16);
        tc_zero_and_release<C>(tlane + j0, tmem_empty, cw);
        TcWitnessStream ws;
        if (FUSE) {
            ws.warp_row0 = stage + (size_t)(warp - 2) * 32 * TcWitnessStream::ROW_U4;
            ws.row0 = ws.warp_row0 + (size_t)lane * TcWitnessStream::ROW_U4;
            ws.W = m * kc.p.W;
            ws.buf_stride = 0;   // single-buffered
            ws.buf = 0;
            ws.fill = 0;
        }
        uint32_t round = 0;
        for (int g = cluster_id; g < num_groups; g += num_clusters) {
            const int ib = g / groups_j, jb = (g % groups_j) * CS + (int)crank;
            const int gi = ib * TC_BM + il;
            for (int pass = 0; pass < passes; pass++) {
                tc_mbar_wait(tmem_full, round & 1, err);
                tc_fence_after();
                if (threadIdx.x == 64) tc_stamp(tl, round, 2);
                // Phase 1 (on the critical path of the next tile): read this warp's columns of all diagonals and carry
                // them into multi-limb integers T = sum_d dg[d] * 2^(8d).
                constexpr int TW = C::SIGNED ? 6 : 18;
                uint32_t T[CW][TW];
#pragma unroll
                for (int pr = 0; pr < CW / 2; pr++) {
                    if (4 * CW != BJ && 2 * pr >= cw) break;
                    uint32_t dg[2][64];
#pragma unroll
                    for (int d = 0; d < D::NDIAG; d++) tc_ld2(tlane + (uint32_t)(BJ * d + j0 + 2 * pr), dg[0][d], dg[1][d]);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // pin the loaded registers behind the wait (the compiler must not read them earlier)
#pragma unroll
                    for (int d = 0; d < D::NDIAG; d++)
#pragma unroll
                        for (int q = 0; q < 2; q++) asm volatile("" : "+r"(dg[q][d]));
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        if constexpr (C::SIGNED) {
                            fr::carry_signed<D::NDIAG>(dg[q], T[2 * pr + q]);
                        } else {
                            // diagonals d = 4g + r sit at whole-word offsets g for fixed r:
                            // T = Y0 + (Y1 << 8) + (Y2 << 16) + (Y3 << 24) with Y_r[g] = dg[4g + r]
                            dg[q][63] = 0;
                            tc_carry_diagonals(dg[q], T[2 * pr + q]);
                        }
                    }
                }
                // the accumulators are free again: zero them and let the MMA warp start the next tile while the
                // Montgomery reductions (and the rescale witnesses) below run
                tc_fence_before();  // order the TMEM reads above before the zeroing stores / the next MMAs
                if (threadIdx.x == 64) tc_stamp(tl, round, 3);
                tc_zero_and_release<C>(tlane + j0, tmem_empty, cw);
                if (threadIdx.x == 64) tc_stamp(tl, round, 4);
                // Phase 2 (overlaps the next tile's MMAs)
                Fr res[CW];
#pragma unroll
                for (int q = 0; q < CW; q++) {
                    if (4 * CW != BJ && q >= cw) break;
                    if constexpr (C::SIGNED) res[q] = fr::signed6_to_mont(T[q]);
                    else res[q] = fr::reduce_wide_acc(T[q]);
                    const int gj = jb * BJ + j0 + q;
                    if (gi < n && gj < m) {
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
                    for (int q = 0; q < CW; q++) {
                        const int gj = jb * BJ + j0 + q;
                        if (gj < m && valid > 0) {  // warp-uniform
                            ws.valid = valid;
                            ws.gwarp = out_wit + ((size_t)row0 * m + gj) * (size_t)kc.p.W;
                            const Fr qv = rs::rescale_element(ws, kc, gi < n ? res[q] : fr::zero());
                            if (gi < n) st_fr(out_q + (size_t)gi * m + gj, qv);
                        }
                    }
                }
                if (threadIdx.x == 64) tc_stamp(tl, round, 5);
                round++;
            }
        }
        // shared memory must outlive every bulk read, and the witness writes must be complete at kernel end
        if (FUSE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) tc_cluster_sync();   // no CTA's shared memory goes away while a peer may still write or signal into it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

using tc_encode_fn = tma_encode_fn;
inline tc_encode_fn tc_encoder() { return tma_encoder(); }

// ---- tensor-pipe micro-benchmark: the measured denominator of the mat-mul roofline --------------------------------
// One CTA per SM, operands resident in shared memory (pseudo-random bytes, the layouts of the real kernel), one elected
// thread issues back-to-back kind::i8 MMAs of the real kernel's shape (M = 128, K = 32, N = NMMA) into TMEM; batches of 32
// instructions are committed to two alternating mbarriers so that the next batch is always queued before the previous
// one is waited for.  Nothing else runs: this is what the pipe delivers when operand delivery and epilogues cost nothing.
template <int NMMA, bool SIGNED>
__global__ void __launch_bounds__(128, 1) tc_peak_kernel(int batches, int* err) {
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t raw = tc_smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = tc_smem_raw + (base - raw);
    const uint32_t s_a = base, s_b = base + TC_A_BYTES, bars = s_b + NMMA * TC_BKB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bars + 16 - base));
    constexpr uint32_t IDESC = (2u << 4) | ((SIGNED ? 1u : 0u) << 7) | ((SIGNED ? 1u : 0u) << 10) |
                               ((uint32_t)(NMMA >> 3) << 17) | ((128u >> 4) << 24);
    for (uint32_t i = threadIdx.x; i < (TC_A_BYTES + NMMA * TC_BKB) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    if (threadIdx.x == 0) {
        tc_mbar_init(bars, 1);
        tc_mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (threadIdx.x == 0) {
        const uint64_t adesc = tc_smem_desc(s_a), bdesc = tc_smem_desc(s_b);
        for (int b = 0; b < batches; b++) {
            if (b >= 2) tc_mbar_wait(bars + 8 * (b & 1), ((b >> 1) - 1) & 1, err);   // batch b-2 has drained
#pragma unroll
            for (int i = 0; i < 32; i++)
                tc_mma_i8(tmem_base + (uint32_t)(i & 1) * (512 - NMMA), adesc + 2u * (i & 3), bdesc + 2u * (i & 3), IDESC, 1u);
            tc_commit(bars + 8 * (b & 1));
        }
        for (int b = batches > 2 ? batches - 2 : 0; b < batches; b++) tc_mbar_wait(bars + 8 * (b & 1), (b >> 1) & 1, err);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int NMMA, bool SIGNED>
int tc_peak_run(h2svd_ctx* ctx, double min_seconds, double* ops_per_s) {
    const size_t smem = TC_A_BYTES + (size_t)NMMA * TC_BKB + 1024 + 64;
    H2SVD_SET_SMEM(ctx, (tc_peak_kernel<NMMA, SIGNED>), smem);
    const int batches = 2048;   // 65536 MMAs per CTA: ~4 ms at N = 256
    cudaEvent_t e0, e1;
    H2SVD_CUDA(cudaEventCreate(&e0));
    H2SVD_CUDA(cudaEventCreate(&e1));
    tc_peak_kernel<NMMA, SIGNED><<<ctx->sm_count, 128, smem, ctx->stream>>>(batches, ctx->d_flag);   // warm-up
    H2SVD_LAUNCH_CHECK(ctx);
    double best = 0, total_ms = 0;
    long launches = 0;
    do {
        H2SVD_CUDA(cudaEventRecord(e0, ctx->stream));
        tc_peak_kernel<NMMA, SIGNED><<<ctx->sm_count, 128, smem, ctx->stream>>>(batches, ctx->d_flag);
        H2SVD_LAUNCH_CHECK(ctx);
        H2SVD_CUDA(cudaEventRecord(e1, ctx->stream));
        H2SVD_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        H2SVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = 2.0 * 128 * NMMA * 32 * 32.0 * batches * ctx->sm_count / (ms * 1e-3);
        if (min_seconds <= 0 && rate > best) best = rate;
        total_ms += ms;
        launches++;
    } while (min_seconds > 0 ? total_ms < min_seconds * 1e3 : launches < 3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // burst: best single launch; sustained: everything issued over >= min_seconds, back to back
    *ops_per_s = min_seconds > 0 ? 2.0 * 128 * NMMA * 32 * 32.0 * batches * ctx->sm_count * launches / (total_ms * 1e-3) : best;
    return H2SVD_OK;
}

size_t tc_align256(size_t x) { return (x + 255) & ~(size_t)255; }

// split kernels + mat-mul kernel of ONE engine on pre-carved plane buffers
template <class C, int CS = 1>
int tc_launch_engine(h2svd_ctx* ctx, tc_encode_fn encode, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m,
                     size_t ldk, uint8_t* a8, uint8_t* b8, int* mode, int run_if_mode, bool split_only, bool mm_only,
                     const rs::RescaleConsts* fuse, Fr* out_q, Fr* out_wit) {
    using D = TcD<C>;
    const int ldk4 = (int)(ldk / 4);
    if (!mm_only) {
        const size_t blocks_a = (n * (size_t)ldk4 + 255) / 256, tiles_k = (ldk + 127) / 128, blocks_b = tiles_k * ((m + 7) / 8);
        if (blocks_a + blocks_b >= (1ull << 31)) {
            set_error("fr_matmul (tensor-core engine): operands too large for the split kernel's grid");
            return H2SVD_EINVAL;
        }
        tc_split_kernel<C><<<(unsigned)(blocks_a + blocks_b), 256, 0, ctx->stream>>>(
            a, reinterpret_cast<uint32_t*>(a8), (int)n, b, reinterpret_cast<uint32_t*>(b8), (int)k, (int)m, ldk4, mode,
            (unsigned)blocks_a, (unsigned)tiles_k);
        H2SVD_LAUNCH_CHECK(ctx);
    }
    if (split_only) return H2SVD_OK;
    CUtensorMap tm_a, tm_b;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)(C::LA * n)};
        const cuuint64_t strides[1] = {(cuuint64_t)ldk};
        const cuuint32_t box[2] = {(cuuint32_t)TC_BKB, (cuuint32_t)(TC_BM / CS)};   // a cluster's CTAs load a slice each
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
        const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)m, (cuuint64_t)C::LB};
        const cuuint64_t strides[2] = {(cuuint64_t)ldk, (cuuint64_t)(m * ldk)};
        const cuuint32_t box[3] = {(cuuint32_t)TC_BKB, (cuuint32_t)C::BJ, (cuuint32_t)C::LB};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&tm_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, b8, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("fr_matmul (tensor-core engine): tensor map for B failed (CUresult %d)", (int)r);
            return H2SVD_ECUDA;
        }
    }
    const int tiles_i = (int)((n + TC_BM - 1) / TC_BM), tiles_j = (int)((m + C::BJ - 1) / C::BJ);
    const long long tiles = (long long)tiles_i * tiles_j;
    if (tiles >= (1LL << 31)) {
        set_error("fr_matmul (tensor-core engine): too many tiles");
        return H2SVD_EINVAL;
    }
    if constexpr (CS > 1) {
        // clusters of CS CTAs along the columns: (tiles_j rounded up to CS) column tiles per row block
        static_assert(TC_BM % (8 * CS) == 0, "a slice is whole swizzle atoms");
        if (fuse) {
            set_error("fr_matmul (tensor-core engine): the fused rescale epilogue has no cluster form");
            return H2SVD_EINVAL;
        }
        const long long groups = (long long)tiles_i * ((tiles_j + CS - 1) / CS);
        const size_t smem = tc_smem_bytes<C, false>();
        auto kern = fr_matmul_tc_kernel<C, false, CS>;
        H2SVD_SET_SMEM(ctx, kern, smem);
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)(ctx->sm_count / CS * CS));
        int max_clusters = 0;   // co-resident clusters (GPC shapes can leave a few SMs out)
        H2SVD_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
        if (max_clusters < 1) {
            set_error("fr_matmul (tensor-core engine): no cluster of %d CTAs fits this device", CS);
            return H2SVD_ECUDA;
        }
        const long long clusters = groups < max_clusters ? groups : max_clusters;
        cfg.gridDim = dim3((unsigned)(clusters * CS));
        static const rs::RescaleConsts none{};
        H2SVD_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_a, tm_b, c, (int)n, (int)k, (int)m, tiles_j,
                                      (int)tiles, ctx->d_flag, (const int*)mode, run_if_mode, ctx->d_timeline, none,
                                      (Fr*)nullptr, (Fr*)nullptr));
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    const int grid = (int)(tiles < ctx->sm_count ? tiles : ctx->sm_count);
    if (fuse) {
        if constexpr (C::BJ == 8) {   // the fused epilogue is instantiated for the 8-column tiles of either engine
            H2SVD_SET_SMEM(ctx, (fr_matmul_tc_kernel<C, true, 1>), (tc_smem_bytes<C, true>()));
            fr_matmul_tc_kernel<C, true, 1><<<grid, TC_THREADS, tc_smem_bytes<C, true>(), ctx->stream>>>(
                tm_a, tm_b, c, (int)n, (int)k, (int)m, tiles_j, (int)tiles, ctx->d_flag, mode, run_if_mode, ctx->d_timeline, *fuse, out_q,
                out_wit);
        } else {
            set_error("fr_matmul (tensor-core engine): fused rescale needs 8-column tiles");
            return H2SVD_EINVAL;
        }
    } else {
        static const rs::RescaleConsts none{};
        H2SVD_SET_SMEM(ctx, (fr_matmul_tc_kernel<C, false, 1>), (tc_smem_bytes<C, false>()));
        fr_matmul_tc_kernel<C, false, 1><<<grid, TC_THREADS, tc_smem_bytes<C, false>(), ctx->stream>>>(
            tm_a, tm_b, c, (int)n, (int)k, (int)m, tiles_j, (int)tiles, ctx->d_flag, mode, run_if_mode, ctx->d_timeline, none, nullptr,
            nullptr);
    }
    H2SVD_LAUNCH_CHECK(ctx);
    (void)D::NMMA;
    return H2SVD_OK;
}

}  // namespace

int launch_microbench_i8(h2svd_ctx* ctx, int kind, double min_seconds, double* ops_per_s) {
    if (!ops_per_s || kind < 0 || kind > 3) {
        set_error("microbench_tensor_i8: bad arguments");
        return H2SVD_EINVAL;
    }
    switch (kind) {   // 0 / 1: the shapes of the two engines; 2, 3: narrower N (the per-instruction floor)
        case 0: return tc_peak_run<TcD<TcFull>::NMMA, false>(ctx, min_seconds, ops_per_s);
        case 1: return tc_peak_run<TcD<TcSmall28>::NMMA, true>(ctx, min_seconds, ops_per_s);
        case 2: return tc_peak_run<TcD<TcSmall16>::NMMA, true>(ctx, min_seconds, ops_per_s);
        default: return tc_peak_run<TcD<TcSmall8>::NMMA, true>(ctx, min_seconds, ops_per_s);
    }
}

bool fr_matmul_tc_supported(size_t n, size_t k, size_t m) {
    // p * n + ib * 128 and the plane sizes are 32-bit TMA coordinates / comfortably below 2^31
    return k >= 1 && n * 32 < (1ull << 31) && m * 32 < (1ull << 31) && k < (1ull << 30);
}

// Tile width of the small-operand engine for one product.  Measured (tools/cluster_bench.py, tools/tc_timeline.py): a tile's
// MMA phase lasts about as long for 8, 16, 24 and 28 columns (~2.5-3 us per 1024 bytes of K: every MMA streams its
// 128 x 32-byte A operand from shared memory in ~128 clocks whatever its N), so the time of a product is the number of
// WAVES of tiles times that, plus the field arithmetic of the last tile (proportional to its columns per warp).  The
// widest tile wins whenever it saves a wave; otherwise the narrowest with the same wave count.
static int tc_small_tile_width(const h2svd_ctx* ctx, size_t n, size_t k, size_t m) {
    const int w = ctx->tune.matmul_small_width;
    if (w == 8 || w == 16 || w == 24 || w == 28) return w;
    const double sms = ctx->sm_count, kblocks = (double)((k + TC_BKB - 1) / TC_BKB);
    const int widths[4] = {8, 16, 24, 28};
    int best = 28;
    double best_t = 1e300;
    for (int i = 0; i < 4; i++) {
        const double bj = widths[i];
        const double tiles = (double)((n + TC_BM - 1) / TC_BM) * (double)((m + widths[i] - 1) / widths[i]);
        const double waves = (double)(long long)((tiles + sms - 1) / sms);
        const double tile = kblocks * (4600.0 + 40.0 * bj) + 700.0;                 // clocks; + accumulator hand-over
        const double t = waves * tile + 2900.0 * (double)(((widths[i] + 3) / 4 + 1) / 2 * 2);   // + the last tile's field arithmetic
        if (t < best_t) {
            best_t = t;
            best = widths[i];
        }
    }
    return best;
}

int launch_fr_matmul_tc(h2svd_ctx* ctx, const Fr* a, const Fr* b, Fr* c, size_t n, size_t k, size_t m,
                        const rs::RescaleConsts* fuse, Fr* out_q, Fr* out_wit) {
    if (n == 0 || m == 0) return H2SVD_OK;
    tc_encode_fn encode = tc_encoder();
    if (!encode) {
        set_error("fr_matmul (tensor-core engine): cuTensorMapEncodeTiled is not available from this driver");
        return H2SVD_ECUDA;
    }
    const bool try_small = ctx->tune.matmul_small != 0;
    const size_t ldk = (k + 15) & ~(size_t)15;  // TMA row pitch: multiple of 16 bytes
    const size_t full_a = tc_align256(TcFull::LA * n * ldk), full_b = tc_align256(TcFull::LB * m * ldk);
    const size_t small_a = tc_align256(TcSmall::LA * n * ldk), small_b = tc_align256(TcSmall::LB * m * ldk);   // LB = 10: the largest
    const size_t need = full_a + full_b + (try_small ? small_a + small_b : 0);
    // the operand workspace is shared with the Karatsuba engine (never both at once)
    H2SVD_TRY(ws_grow(ctx, &ctx->kara_ws, &ctx->kara_ws_bytes, need));
    uint8_t* fa8 = reinterpret_cast<uint8_t*>(ctx->kara_ws);
    uint8_t* fb8 = fa8 + full_a;
    uint8_t* sa8 = fb8 + full_b;
    uint8_t* sb8 = sa8 + small_a;
    if (!try_small) {
        ctx->last_engine = 2;
        return tc_launch_engine<TcFull>(ctx, encode, a, b, c, n, k, m, ldk, fa8, fb8, nullptr, 0, false, false, fuse, out_q,
                                        out_wit);
    }
    // Both engines are enqueued; *d_mode (0 = every operand is a small signed integer, else 1), written by the
    // small-operand split kernel, decides on the device which of them does the work.
    ctx->last_engine = 3;
    H2SVD_CUDA(cudaMemsetAsync(ctx->d_mode, 0, sizeof(int), ctx->stream));
    const int width = fuse ? 8 : tc_small_tile_width(ctx, n, k, m);
    // clusters of 2 CTAs sharing each A-plane stage by TMA multicast: byte-identical, measured to change nothing
    // (tuning switch "matmul_cluster", off) -- the tile is bound by the MMAs' own operand reads, not by L2 delivery
    const int cluster = fuse ? 1 : ctx->tune.matmul_cluster;
    auto small = [&](bool split_only, bool mm_only) -> int {
#define H2SVD_SMALL_ARGS ctx, encode, a, b, c, n, k, m, ldk, sa8, sb8, ctx->d_mode, 0, split_only, mm_only
        if (cluster == 2 && !split_only) {
            switch (width) {
                case 8: return tc_launch_engine<TcSmall8, 2>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
                case 16: return tc_launch_engine<TcSmall16, 2>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
                case 24: return tc_launch_engine<TcSmall, 2>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
                default: return tc_launch_engine<TcSmall28, 2>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
            }
        }
        switch (width) {
            case 8: return tc_launch_engine<TcSmall8>(H2SVD_SMALL_ARGS, fuse, out_q, out_wit);
            case 16: return tc_launch_engine<TcSmall16>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
            case 24: return tc_launch_engine<TcSmall>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
            default: return tc_launch_engine<TcSmall28>(H2SVD_SMALL_ARGS, nullptr, nullptr, nullptr);
        }
#undef H2SVD_SMALL_ARGS
    };
    H2SVD_TRY(small(true, false));
    H2SVD_TRY(tc_launch_engine<TcFull>(ctx, encode, a, b, c, n, k, m, ldk, fa8, fb8, ctx->d_mode, 1, true, false, nullptr,
                                       nullptr, nullptr));
    // (Finishing a partial last wave of wide tiles with 128 x 8 tiles was measured and dropped: narrow tiles take as long.)
    H2SVD_TRY(small(false, true));
    return tc_launch_engine<TcFull>(ctx, encode, a, b, c, n, k, m, ldk, fa8, fb8, ctx->d_mode, 1, false, true, fuse, out_q,
                                    out_wit);
}

}  // namespace h2svd
