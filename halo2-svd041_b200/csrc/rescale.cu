// K4: rescale witnesses -- the value computation of ZkMatrix::rescale_matrix
// (reference src/matrix/mod.rs:354-375), i.e. FixedPointChip041::signed_div_scale per element
// (third-party; cell model in SURVEY.md A.4/A.5):
//
//   a_shift = a + 2^S                                   gate.add           1 Witness
//   (div, rem) = div_mod_floor(a_shift, 2^P)            range.div_mod      2 Witness (rem, div)
//   check_big_less_than_safe(div, 2^(A-P) + 1)          4*n_d Witness
//   check_big_less_than_safe(rem, 2^P)                  4*n_r Witness
//   q = div - 2^(S-P)                                   gate.sub           1 Witness
//
// where each check_big_less_than_safe(x, B) with n = ceil(B.bits()/lb), bits = n*lb emits
//   range_check(x, bits):        l0, l1, s1, l2, s2, ... (2n-1 values; nothing when n == 1)
//   check_less_than(x, B, bits): x + 2^bits - B,  x + 2^bits
//   range_check(x + 2^bits - B, bits): 2n-1 values again.
//
// The kernel is bound by the HBM write of W 32-byte witnesses per element (1920 B at P=63, lb=19), so
// the design minimises integer work per witness and makes every store a full-line bulk write:
//  * Montgomery form is linear, so a limb l at position i enters the running sum as
//    M(l * 2^(lb*i)) = l * c_i * 2^-32 mod r with c_i = 2^(lb*i) * 2^288 mod r precomputed on the host:
//    ONE single-limb Montgomery step (16 IMAD.WIDE, fr_fast.cuh) instead of a full 8x8 conversion
//    (~264 IMAD) per witness; running sums are modular adds; x + 2^bits, x + 2^bits - B and q are modular
//    adds/subs of M(x) with host-precomputed Montgomery constants.
//  * one thread per element; each lane stages its witnesses in its own shared-memory row
//    (double-buffered, 16-byte skew => conflict-free) and ships every CH witnesses with one TMA bulk
//    store (cp.async.bulk.global.shared::cta): all global writes are contiguous CH*32-byte bursts
//    issued by the copy engine, no thread ever issues a 32-byte scattered store.
#include "common.cuh"
#include "fr_fast.cuh"
#include "rescale_dev.cuh"

namespace h2svd {

namespace {

using namespace rs;

// check_abs_less_than / check_mat_diff (reference src/matrix/mod.rs:425-459)
struct AbsLtConsts {
    int n, W, with_diff;
    Fr m_add;                                // M(bnd - 1)
    Fr i_pow, i_bound, m_pow, m_bound;       // 2^(n*lb), 2*bnd - 1
    LimbConsts lc;
};
// RangeChip::range_check(x, range_bits)
struct RangeConsts {
    int n, rem, W;
    Fr c_shift;                              // 2^(lb - rem) * 2^288 mod r (rem > 1)
    Fr m_shift;                              // M(2^(lb - rem))
    LimbConsts lc;
};

// ---------------------------------------------------------------------------------------------------
// generic fallback (any lb >= 1): full Montgomery conversion per witness, scattered 32-byte stores.
__device__ __forceinline__ Fr* emit(Fr* w, const Fr& x_int) {
    st_fr_cs(w, fr::to_mont(x_int));
    return w + 1;
}
__device__ __forceinline__ Fr* emit_range_check(Fr* w, const Fr& x, int n, int lb) {
    if (n == 1) return w;
    for (int i = 0; i < n; i++) {
        w = emit(w, fr::low_bits(fr::shr(x, lb * i), lb));
        if (i >= 1) w = emit(w, fr::low_bits(x, lb * (i + 1)));
    }
    return w;
}
__device__ __forceinline__ Fr* emit_cbls(Fr* w, const Fr& x, const Fr& bound, int n, int lb) {
    w = emit_range_check(w, x, n, lb);
    const Fr xp = fr::add(x, fr::pow2(n * lb));  // x + 2^bits      (mod r)
    const Fr chk = fr::sub(xp, bound);           // x + 2^bits - B  (mod r)
    w = emit(w, chk);
    w = emit(w, xp);
    return emit_range_check(w, chk, n, lb);
}
__global__ void __launch_bounds__(128)
rescale_generic_kernel(const Fr* __restrict__ cs, Fr* __restrict__ out_q, Fr* __restrict__ out_wit, size_t count,
                       RescaleParams p) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const Fr a = fr::from_mont(ldg_fr(cs + e));            // canonical integer
    const Fr ash = fr::add(a, fr::pow2(p.S));              // gate.add(a, Constant(2^S))
    const Fr div = fr::shr(ash, p.P);                      // div_mod_floor by 2^P
    const Fr rem = fr::low_bits(ash, p.P);
    Fr bound_d = fr::pow2(p.A - p.P);
    bound_d.l[0] |= 1u;                                    // 2^A / 2^P + 1   (A > P)
    const Fr bound_r = fr::pow2(p.P);
    Fr* w = out_wit + e * (size_t)p.W;
    w = emit(w, ash);
    w = emit(w, rem);
    w = emit(w, div);
    w = emit_cbls(w, div, bound_d, p.n_d, p.lb);
    w = emit_cbls(w, rem, bound_r, p.n_r, p.lb);
    const Fr q = fr::to_mont(fr::sub(div, fr::pow2(p.S - p.P)));  // gate.sub(div, Constant(2^(S-P)))
    st_fr_cs(w, q);
    st_fr(out_q + e, q);
}

// ---------------------------------------------------------------------------------------------------
// staged kernel
constexpr int RS_THREADS = 128;
#ifndef RS_CH_CFG
#define RS_CH_CFG 8
#endif
#ifndef RS_NBUF_CFG
#define RS_NBUF_CFG 2
#endif
constexpr int RS_CH = RS_CH_CFG;              // witnesses per bulk store (256 B)
constexpr int RS_NBUF = RS_NBUF_CFG;          // staging rows per lane
constexpr int RS_ROW_U4 = RS_CH * 2 + 1;      // staging row in 16-byte units: 256 B + 16 B skew
constexpr size_t RS_SMEM = (size_t)RS_THREADS * RS_NBUF * RS_ROW_U4 * sizeof(uint4);
constexpr int RS_CTAS_PER_SM = (int)((227 * 1024) / (RS_SMEM + 1024)) > 0 ? (int)((227 * 1024) / (RS_SMEM + 1024)) : 1;
using WitnessStream = WitnessStreamT<RS_CH, RS_NBUF>;   // 256-byte bursts, double-buffered rows
// the rescale kernel itself is instantiated for several burst sizes (tuning switch "rescale_ch"): fewer witnesses per
// burst = a smaller staging area = more resident CTAs per SM (the kernel is latency-, not bandwidth-limited)
template <int CH, int NBUF = RS_NBUF, int MAXCTA = 4>
struct RsCfg {
    static constexpr int ROW_U4 = CH * 2 + 1;
    static constexpr size_t SMEM = (size_t)RS_THREADS * NBUF * ROW_U4 * sizeof(uint4);
    static constexpr int CTAS_PER_SM_SMEM = (int)((227 * 1024) / (SMEM + 1024));
    static constexpr int CTAS_PER_SM = CTAS_PER_SM_SMEM < MAXCTA ? CTAS_PER_SM_SMEM : MAXCTA;   // 128 threads x 120 registers: 4 by registers (5: 96)
};

template <int CH, int NBUF, int MAXCTA, bool FAST>
__global__ void __launch_bounds__(RS_THREADS, RsCfg<CH, NBUF, MAXCTA>::CTAS_PER_SM)
rescale_kernel(const Fr* __restrict__ cs, Fr* __restrict__ out_q, Fr* __restrict__ out_wit, size_t count,
               const __grid_constant__ RescaleConsts k) {
    extern __shared__ __align__(16) uint4 rs_stage[];
    const int lane = threadIdx.x & 31;
    WitnessStreamT<CH, NBUF> ws;
    ws.row0 = rs_stage + (size_t)threadIdx.x * RsCfg<CH, NBUF>::ROW_U4;
    ws.warp_row0 = rs_stage + (size_t)(threadIdx.x - lane) * RsCfg<CH, NBUF>::ROW_U4;
    ws.buf_stride = RS_THREADS * RsCfg<CH, NBUF>::ROW_U4;
    ws.W = k.p.W;
    ws.buf = 0;
    ws.fill = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: lanes past the end recompute the last element and ship nothing
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < count; e0 += stride) {
        const size_t left = count - e0;
        ws.valid = left < 32 ? (int)left : 32;
        ws.gwarp = out_wit + e0 * (size_t)k.p.W;
        const bool live = lane < ws.valid;
        const size_t e = live ? e0 + lane : count - 1;
        const Fr am = ldg_fr(cs + e);
#ifdef RS_STORE_ONLY   // experiment: the store path alone (W copies of the input instead of the witnesses)
        for (int w = 0; w < k.p.W; w++) ws.put(am);
        ws.flush();
        const Fr q = am;
#else
        const Fr q = rescale_element<FAST>(ws, k, am);
#endif
        if (live) st_fr(out_q + e, q);
    }
    // shared memory must outlive every bulk read, and the writes must be complete at kernel end.  Bulk async-groups
    // belong to the thread that committed them and elect.sync need not pick lane 0: every lane waits (a no-op for lanes
    // that never committed a group).
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The same kernel with 256-byte-aligned TMA tensor stores for the witness stream (WitnessStreamTma; W % 8 == 4).
constexpr int RT_THREADS = 128;
using RtStream = WitnessStreamTma;
constexpr size_t RT_SMEM = (size_t)(RT_THREADS / 32) * RtStream::WARP_BYTES + 1024;   // + alignment slack: 73 KB, 3 CTAs per SM
template <bool FAST, int CTAS>
__global__ void __launch_bounds__(RT_THREADS, CTAS)
rescale_tma_kernel(const Fr* __restrict__ cs, Fr* __restrict__ out_q, size_t count, const __grid_constant__ CUtensorMap map_even,
                   const __grid_constant__ CUtensorMap map_odd, const __grid_constant__ CUtensorMap map_head,
                   const __grid_constant__ RescaleConsts k) {
    extern __shared__ uint8_t rt_stage_raw[];
    const uint32_t raw = smem_addr(rt_stage_raw);
    uint8_t* stage = rt_stage_raw + (((raw + 1023u) & ~1023u) - raw);   // the swizzle pattern repeats every 1024 bytes
    RtStream ws;
    ws.init(stage + (size_t)(threadIdx.x >> 5) * RtStream::WARP_BYTES, &map_even, &map_odd, &map_head);
    const int lane = ws.lane;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: lanes past the end recompute the last element; their rows are outside the tensors
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < count; e0 += stride) {
        ws.begin((int)e0);
        const bool live = e0 + lane < count;
        const size_t e = live ? e0 + lane : count - 1;
        const Fr am = ldg_fr(cs + e);
#ifdef RS_STORE_ONLY
        for (int w = 0; w < k.p.W; w++) ws.put(am);
        ws.flush();
        const Fr q = am;
#else
        const Fr q = rescale_element<FAST>(ws, k, am);
#endif
        if (live) st_fr(out_q + e, q);
    }
    // shared memory must outlive every TMA read, and the writes must be complete at kernel end (every lane: the groups
    // belong to whichever lane elect.sync picked)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The same kernel with plain coalesced stores for the witness stream (WitnessStreamStg), one staging row per lane.
constexpr int RG_THREADS = 128;
constexpr int RG_CH = 8;
using RgStream = WitnessStreamStg<RG_CH>;
constexpr size_t RG_SMEM = (size_t)RG_THREADS * RgStream::ROW_U4 * sizeof(uint4);   // 34 KB
__global__ void __launch_bounds__(RG_THREADS, 4)
rescale_stg_kernel(const Fr* __restrict__ cs, Fr* __restrict__ out_q, Fr* __restrict__ out_wit, size_t count,
                   const __grid_constant__ RescaleConsts k) {
    extern __shared__ __align__(16) uint4 rs_stage[];
    const int lane = threadIdx.x & 31;
    RgStream ws;
    ws.row0 = rs_stage + (size_t)threadIdx.x * RgStream::ROW_U4;
    ws.warp_row0 = rs_stage + (size_t)(threadIdx.x - lane) * RgStream::ROW_U4;
    ws.W = k.p.W;
    ws.lane = lane;
    ws.fill = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < count; e0 += stride) {
        const size_t left = count - e0;
        ws.valid = left < 32 ? (int)left : 32;
        ws.gwarp = out_wit + e0 * (size_t)k.p.W;
        const bool live = lane < ws.valid;
        const size_t e = live ? e0 + lane : count - 1;
        const Fr am = ldg_fr(cs + e);
        const Fr q = rescale_element(ws, k, am);
        if (live) st_fr(out_q + e, q);
    }
}

// check_abs_less_than(x, bnd) (reference src/matrix/mod.rs:425-437), optionally of a difference x - y
// (check_mat_diff :441-459): witnesses [x - y]?, t = d + (bnd - 1), check_big_less_than_safe(t, 2*bnd - 1).
template <bool FAST>
__global__ void __launch_bounds__(RS_THREADS)
abs_less_than_kernel(const Fr* __restrict__ x, const Fr* __restrict__ y, Fr* __restrict__ out_wit, size_t count,
                     const __grid_constant__ AbsLtConsts k) {
    extern __shared__ __align__(16) uint4 rs_stage[];
    const int lane = threadIdx.x & 31;
    WitnessStream ws;
    ws.row0 = rs_stage + (size_t)threadIdx.x * RS_ROW_U4;
    ws.warp_row0 = rs_stage + (size_t)(threadIdx.x - lane) * RS_ROW_U4;
    ws.buf_stride = RS_THREADS * RS_ROW_U4;
    ws.W = k.W;
    ws.buf = 0;
    ws.fill = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < count; e0 += stride) {
        const size_t left = count - e0;
        ws.valid = left < 32 ? (int)left : 32;
        ws.gwarp = out_wit + e0 * (size_t)k.W;
        const size_t e = lane < ws.valid ? e0 + lane : count - 1;
        Fr dm = ldg_fr(x + e);
        if (k.with_diff) {
            dm = fr::sub_fast(dm, ldg_fr(y + e));   // gate.sub(a, b) -> Witness(a - b)
            ws.put(dm);
        }
        const Fr tm = fr::add_fast(dm, k.m_add);     // gate.add(x, Constant(bnd - 1))
        ws.put(tm);
        stream_cbls<FAST>(ws, k.lc, fr::mont_reduce_fast(tm), tm, k.n, k.i_pow, k.i_bound, k.m_pow, k.m_bound);
        ws.flush();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every lane: see rescale_kernel
}

// RangeChip::range_check(x, range_bits) with n = ceil(range_bits / lb) limbs: limbs + running sums (none when
// n == 1), plus last_limb * 2^(lb - rem) when range_bits % lb = rem > 1 (ZkVector::entries_less_than, :185-197).
template <bool FAST>
__global__ void __launch_bounds__(RS_THREADS)
range_check_kernel(const Fr* __restrict__ x, Fr* __restrict__ out_wit, size_t count, const __grid_constant__ RangeConsts k) {
    extern __shared__ __align__(16) uint4 rs_stage[];
    const int lane = threadIdx.x & 31;
    WitnessStream ws;
    ws.row0 = rs_stage + (size_t)threadIdx.x * RS_ROW_U4;
    ws.warp_row0 = rs_stage + (size_t)(threadIdx.x - lane) * RS_ROW_U4;
    ws.buf_stride = RS_THREADS * RS_ROW_U4;
    ws.W = k.W;
    ws.buf = 0;
    ws.fill = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e0 = (size_t)blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < count; e0 += stride) {
        const size_t left = count - e0;
        ws.valid = left < 32 ? (int)left : 32;
        ws.gwarp = out_wit + e0 * (size_t)k.W;
        const size_t e = lane < ws.valid ? e0 + lane : count - 1;
        const Fr xm = ldg_fr(x + e);
        const Fr xi = fr::mont_reduce_fast(xm);
        stream_range_check<FAST>(ws, k.lc, xi, k.n);
        if (k.rem > 1) {
            if (k.n == 1) {
                ws.put(fr::mont_mul_fast(xm, k.m_shift));   // gate.mul(x, 2^(lb-rem)): x is the only "limb", at full width
            } else {
                Fr t = xi;                                   // last limb: bits [lb*(n-1), lb*n) of x
                for (int i = 0; i < k.n - 1; i++) t = shr_small(t, k.lc.lb);
                ws.put(fr::mont_mul_small(t.l[0] & k.lc.lb_mask, k.c_shift));
            }
        }
        ws.flush();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every lane: see rescale_kernel
}

// mat_times_diag_mat (reference :610-627): out[i][j] = a[i][j] * v[j], j < cols_v <= lda
__global__ void mat_times_diag_kernel(const Fr* __restrict__ a, const Fr* __restrict__ v, Fr* __restrict__ out, size_t rows,
                                      size_t lda, size_t cols_v) {
    const size_t total = rows * cols_v;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const size_t i = idx / cols_v, j = idx - i * cols_v;
        st_fr(out + idx, fr::mont_mul_fast(ldg_fr(a + i * lda + j), ldg_fr(v + j)));
    }
}

}  // namespace

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

int rescale_params(int P, int lb, int S, int A, int* n_d, int* n_r) {
    if (S < 0) S = 3 * P;
    if (A < 0) A = 4 * P;
    if (P < 1 || P > 63 || lb < 1 || lb > 32 || S < P || S > 252 || A <= P) return -1;
    const int nd = ceil_div(A - P + 1, lb), nr = ceil_div(P + 1, lb);
    if (nd * lb > 253 || nr * lb > 253) return -1;
    if (n_d) *n_d = nd;
    if (n_r) *n_r = nr;
    // per check_big_less_than_safe: range_check (2n-1) + chk, xp + range_check (2n-1) = 4n; only chk, xp when n == 1
    return 4 + (nd >= 2 ? 4 * nd : 2) + (nr >= 2 ? 4 * nr : 2);
}

// floor(x * 2^64 / r) for canonical x < r: 64 steps of shift-compare-subtract
static uint64_t frac64_of_r(const Fr& x) {
    uint32_t m[8], rem[8];
    for (int i = 0; i < 8; i++) {
        m[i] = fr::modulus(i);
        rem[i] = x.l[i];
    }
    uint64_t q = 0;
    for (int b = 63; b >= 0; b--) {
        uint32_t carry = 0;   // rem <<= 1 (rem < r < 2^254: no overflow)
        for (int i = 0; i < 8; i++) {
            const uint32_t nc = rem[i] >> 31;
            rem[i] = (rem[i] << 1) | carry;
            carry = nc;
        }
        bool ge = true;
        for (int i = 7; i >= 0; i--)
            if (rem[i] != m[i]) {
                ge = rem[i] > m[i];
                break;
            }
        if (ge) {
            fr::sub_n<8>(rem, rem, m);
            q |= 1ull << b;
        }
    }
    return q;
}
static void fill_limb_consts(LimbConsts& lc, int lb, int npos) {
    lc.lb = lb;
    lc.lb_mask = lb == 32 ? 0xffffffffu : ((1u << lb) - 1u);
    // sum of npos limbs < npos * 2^lb must stay below 2^32 (fr::SmallSum)
    lc.fast_sums = ((unsigned long long)npos << lb) <= (1ull << 32) ? 1 : 0;
    const Fr m_2_32 = fr::to_mont(fr::pow2(32));
    for (int i = 0; i < MAX_POS; i++) {
        lc.rho[i] = i < npos ? fr::to_mont(fr::pow2(lb * i)) : fr::zero();
        lc.c[i] = i < npos ? fr::mont_mul(lc.rho[i], m_2_32) : fr::zero();
        const uint64_t phi = frac64_of_r(lc.rho[i]);
        lc.phi_lo[i] = (uint32_t)phi;
        lc.phi_hi[i] = (uint32_t)(phi >> 32);
    }
}
static int bit_length(const Fr& x) {
    for (int i = 7; i >= 0; i--)
        if (x.l[i]) return 32 * i + (32 - __builtin_clz(x.l[i]));
    return 0;
}

int make_rescale_consts(int P, int lb, int S, int A, rs::RescaleConsts* out) {
    if (S < 0) S = 3 * P;
    if (A < 0) A = 4 * P;
    RescaleConsts& k = *out;
    RescaleParams& p = k.p;
    p.P = P; p.lb = lb; p.S = S; p.A = A;
    p.W = rescale_params(P, lb, S, A, &p.n_d, &p.n_r);
    if (p.W < 0 || p.n_d > MAX_POS || p.n_r > MAX_POS) return -1;
    // constants of this configuration (fr.cuh is host-callable)
    fill_limb_consts(k.lc, lb, p.n_d > p.n_r ? p.n_d : p.n_r);
    k.i_2S = fr::pow2(S);
    k.i_pow_d = fr::pow2(p.n_d * lb);
    k.i_bound_d = fr::pow2(A - P);
    k.i_bound_d.l[0] |= 1u;  // 2^A / 2^P + 1   (A > P)
    k.i_pow_r = fr::pow2(p.n_r * lb);
    k.i_bound_r = fr::pow2(P);
    k.m_2S = fr::to_mont(k.i_2S);
    k.m_2SP = fr::to_mont(fr::pow2(S - P));
    k.m_pow_d = fr::to_mont(k.i_pow_d);
    k.m_bound_d = fr::to_mont(k.i_bound_d);
    k.m_pow_r = fr::to_mont(k.i_pow_r);
    k.m_bound_r = fr::to_mont(k.i_bound_r);
    return p.W;
}

template <int CH, int NBUF = RS_NBUF, int MAXCTA = 4, bool FAST = false>
static int launch_rescale_ch(h2svd_ctx* ctx, const Fr* cs, size_t count, const rs::RescaleConsts& k, Fr* out_q, Fr* out_wit) {
    using cfg = RsCfg<CH, NBUF, MAXCTA>;
    H2SVD_SET_SMEM(ctx, (rescale_kernel<CH, NBUF, MAXCTA, FAST>), cfg::SMEM);
    size_t blocks = (count + RS_THREADS - 1) / RS_THREADS;
    const size_t cap = (size_t)ctx->sm_count * cfg::CTAS_PER_SM;  // resident CTAs, grid-stride beyond
    if (blocks > cap) blocks = cap;
    rescale_kernel<CH, NBUF, MAXCTA, FAST><<<(unsigned)blocks, RS_THREADS, cfg::SMEM, ctx->stream>>>(cs, out_q, out_wit, count, k);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_rescale(h2svd_ctx* ctx, const Fr* cs, size_t count, int P, int lb, int S, int A, Fr* out_q,
                   Fr* out_wit) {
    if (S < 0) S = 3 * P;
    if (A < 0) A = 4 * P;
    RescaleParams p;
    p.P = P; p.lb = lb; p.S = S; p.A = A;
    p.W = rescale_params(P, lb, S, A, &p.n_d, &p.n_r);
    if (p.W < 0) {
        set_error("rescale: parameters out of range (P=%d lb=%d S=%d A=%d)", P, lb, S, A);
        return H2SVD_EINVAL;
    }
    if (count == 0) return H2SVD_OK;
    if (p.n_d > MAX_POS || p.n_r > MAX_POS || ctx->tune.rescale_generic) {
        rescale_generic_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(cs, out_q, out_wit, count, p);
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    // the constants of a configuration cost ~50 us of host arithmetic: keep the last one
    static thread_local struct { int P, lb, S, A; bool ok; RescaleConsts k; } cache = {0, 0, 0, 0, false, {}};
    if (!cache.ok || cache.P != P || cache.lb != lb || cache.S != S || cache.A != A) {
        make_rescale_consts(P, lb, S, A, &cache.k);
        cache.P = P; cache.lb = lb; cache.S = S; cache.A = A;
        cache.ok = true;
    }
    const RescaleConsts& k = cache.k;
    if (ctx->tune.rescale_store == 2) {
        H2SVD_SET_SMEM(ctx, rescale_stg_kernel, RG_SMEM);
        size_t blocks = (count + RG_THREADS - 1) / RG_THREADS;
        const size_t cap = (size_t)ctx->sm_count * 4;
        if (blocks > cap) blocks = cap;
        rescale_stg_kernel<<<(unsigned)blocks, RG_THREADS, RG_SMEM, ctx->stream>>>(cs, out_q, out_wit, count, k);
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    const bool fast = k.lc.fast_sums && ctx->tune.rescale_fast_sums != 0;
    if ((ctx->tune.rescale_store == 0 || ctx->tune.rescale_store == 1) && ctx->tune.rescale_ch == 8 && RtStream::usable(out_wit, p.W, count)) {
        // the even and the odd stripes as two 3-D tensors [128 B][halves][elements / 2]; box = 128 B x 2 halves x 16 rows,
        // plus a one-half box for the first 128 bytes of the odd stripes
        tma_encode_fn encode = tma_encoder();
        if (encode) {
            CUtensorMap maps[3];
            bool ok = true;
            for (int c = 0; c < 3 && ok; c++) {
                const int odd = c >= 1;
                const cuuint64_t rows = odd ? count / 2 : (count + 1) / 2;
                const cuuint64_t dims[3] = {128, (cuuint64_t)(p.W / 4), rows};
                const cuuint64_t strides[2] = {128, (cuuint64_t)p.W * 64};
                const cuuint32_t box[3] = {128, c == 2 ? 1u : 2u, 16};
                const cuuint32_t estr[3] = {1, 1, 1};
                ok = encode(&maps[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, reinterpret_cast<uint8_t*>(out_wit) + (size_t)odd * p.W * 32, dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            if (ok) {
                size_t blocks = (count + RT_THREADS - 1) / RT_THREADS;
                const int ctas = ctx->tune.rescale_ctas == 2 ? 2 : 3;
                const size_t cap = (size_t)ctx->sm_count * ctas;
                if (blocks > cap) blocks = cap;
#define H2SVD_RT_LAUNCH(F, C)                                                                                               \
    do {                                                                                                                    \
        H2SVD_SET_SMEM(ctx, (rescale_tma_kernel<F, C>), RT_SMEM);                                                           \
        rescale_tma_kernel<F, C><<<(unsigned)blocks, RT_THREADS, RT_SMEM, ctx->stream>>>(cs, out_q, count, maps[0], maps[1], \
                                                                                         maps[2], k);                       \
    } while (0)
                if (fast && ctas == 2) H2SVD_RT_LAUNCH(true, 2);
                else if (fast) H2SVD_RT_LAUNCH(true, 3);
                else if (ctas == 2) H2SVD_RT_LAUNCH(false, 2);
                else H2SVD_RT_LAUNCH(false, 3);
#undef H2SVD_RT_LAUNCH
                H2SVD_LAUNCH_CHECK(ctx);
                return H2SVD_OK;
            }
        }
    }
    // Burst sizes / residencies measured on the N = 1024 rescale: 8 witnesses per burst, two staging
    // rows per lane, 3 CTAs per SM is the fastest; 4- and 6-witness bursts remain as tuning switches.
    switch (ctx->tune.rescale_ch) {
        case 4: return launch_rescale_ch<4>(ctx, cs, count, k, out_q, out_wit);
        case 6: return launch_rescale_ch<6>(ctx, cs, count, k, out_q, out_wit);
        default:
            return fast ? launch_rescale_ch<8, RS_NBUF, 4, true>(ctx, cs, count, k, out_q, out_wit)
                        : launch_rescale_ch<8>(ctx, cs, count, k, out_q, out_wit);
    }
}

// W of check_abs_less_than for bound `bnd` (canonical integer): [diff] + t + cbls(2*bnd - 1)
int abs_less_than_params(const Fr& bnd, int lb, int with_diff, int* n_out, Fr* bound_out) {
    if (lb < 1 || lb > 32 || fr::is_zero(bnd) || bit_length(bnd) > 250) return -1;
    Fr two_bnd, one = fr::zero();
    one.l[0] = 1u;
    fr::add_n<8>(two_bnd.l, bnd.l, bnd.l);
    Fr bound;
    fr::sub_n<8>(bound.l, two_bnd.l, one.l);  // 2*bnd - 1
    const int n = ceil_div(bit_length(bound), lb);
    if (n * lb > 253 || n > MAX_POS) return -1;
    if (n_out) *n_out = n;
    if (bound_out) *bound_out = bound;
    return (with_diff ? 1 : 0) + 1 + 2 + (n >= 2 ? 2 * (2 * n - 1) : 0);
}

int launch_abs_less_than(h2svd_ctx* ctx, const Fr* x, const Fr* y, size_t count, const Fr& bnd, int lb, Fr* out_wit) {
    AbsLtConsts k;
    Fr bound;
    k.with_diff = y != nullptr;
    k.W = abs_less_than_params(bnd, lb, k.with_diff, &k.n, &bound);
    if (k.W < 0) {
        set_error("abs_less_than: parameters out of range");
        return H2SVD_EINVAL;
    }
    if (count == 0) return H2SVD_OK;
    fill_limb_consts(k.lc, lb, k.n);
    Fr one = fr::zero();
    one.l[0] = 1u;
    Fr bm1;
    fr::sub_n<8>(bm1.l, bnd.l, one.l);
    k.m_add = fr::to_mont(bm1);
    k.i_pow = fr::pow2(k.n * lb);
    k.i_bound = bound;
    k.m_pow = fr::to_mont(k.i_pow);
    k.m_bound = fr::to_mont(bound);
    size_t blocks = (count + RS_THREADS - 1) / RS_THREADS;
    const size_t cap = (size_t)ctx->sm_count * RS_CTAS_PER_SM;
    if (blocks > cap) blocks = cap;
    if (k.lc.fast_sums && ctx->tune.rescale_fast_sums != 0) {
        H2SVD_SET_SMEM(ctx, abs_less_than_kernel<true>, RS_SMEM);
        abs_less_than_kernel<true><<<(unsigned)blocks, RS_THREADS, RS_SMEM, ctx->stream>>>(x, y, out_wit, count, k);
    } else {
        H2SVD_SET_SMEM(ctx, abs_less_than_kernel<false>, RS_SMEM);
        abs_less_than_kernel<false><<<(unsigned)blocks, RS_THREADS, RS_SMEM, ctx->stream>>>(x, y, out_wit, count, k);
    }
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int range_check_params(int range_bits, int lb, int* n_out, int* rem_out) {
    if (lb < 1 || lb > 32 || range_bits < 1 || range_bits > 253) return -1;
    const int n = ceil_div(range_bits, lb), rem = range_bits % lb;
    if (n > MAX_POS) return -1;
    if (n_out) *n_out = n;
    if (rem_out) *rem_out = rem;
    return (n >= 2 ? 2 * n - 1 : 0) + (rem > 1 ? 1 : 0);
}

int launch_range_check(h2svd_ctx* ctx, const Fr* x, size_t count, int range_bits, int lb, Fr* out_wit) {
    RangeConsts k;
    k.W = range_check_params(range_bits, lb, &k.n, &k.rem);
    if (k.W < 0) {
        set_error("range_check: parameters out of range");
        return H2SVD_EINVAL;
    }
    if (count == 0 || k.W == 0) return H2SVD_OK;
    fill_limb_consts(k.lc, lb, k.n);
    k.m_shift = k.rem > 1 ? fr::to_mont(fr::pow2(lb - k.rem)) : fr::zero();
    k.c_shift = k.rem > 1 ? fr::mont_mul(k.m_shift, fr::to_mont(fr::pow2(32))) : fr::zero();
    size_t blocks = (count + RS_THREADS - 1) / RS_THREADS;
    const size_t cap = (size_t)ctx->sm_count * RS_CTAS_PER_SM;
    if (blocks > cap) blocks = cap;
    if (k.lc.fast_sums && ctx->tune.rescale_fast_sums != 0) {
        H2SVD_SET_SMEM(ctx, range_check_kernel<true>, RS_SMEM);
        range_check_kernel<true><<<(unsigned)blocks, RS_THREADS, RS_SMEM, ctx->stream>>>(x, out_wit, count, k);
    } else {
        H2SVD_SET_SMEM(ctx, range_check_kernel<false>, RS_SMEM);
        range_check_kernel<false><<<(unsigned)blocks, RS_THREADS, RS_SMEM, ctx->stream>>>(x, out_wit, count, k);
    }
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_mat_times_diag(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t lda, size_t cols_v, Fr* out) {
    if (rows == 0 || cols_v == 0) return H2SVD_OK;
    const size_t total = rows * cols_v;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    mat_times_diag_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, v, out, rows, lda, cols_v);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd

