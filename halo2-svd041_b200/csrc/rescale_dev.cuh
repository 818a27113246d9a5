// Device-side pieces of the rescale witness generator (K4) shared by rescale.cu (stand-alone kernels) and
// matmul_tc.cu (the mat-mul epilogue that emits the rescale witnesses of each C element as it is produced).
// See rescale.cu for the cell model (FixedPointChip041::signed_div_scale, reference src/matrix/mod.rs:354-375).
#pragma once
#include "common.cuh"
#include "fr_fast.cuh"
#include "tma_util.cuh"

namespace h2svd {
namespace rs {

constexpr int MAX_POS = 32;  // limb positions the staged kernel supports (n_d, n_r <= 32)

struct RescaleParams {
    int P, lb, S, A, n_d, n_r, W;
};

// Host-precomputed constants of one (P, lb, S, A) configuration, passed by value (constant bank).
// limb-decomposition constants shared by every range-check style kernel
struct LimbConsts {
    int lb;
    uint32_t lb_mask;
    int fast_sums;                                // npos * 2^lb <= 2^32: the SmallSum running sums (fr_fast.cuh) are exact
    Fr c[MAX_POS];                                // c[i] = 2^(lb*i) * 2^288 mod r
    Fr rho[MAX_POS];                              // rho[i] = 2^(lb*i) * 2^256 mod r, canonical integer = M(2^(lb*i))
    uint32_t phi_lo[MAX_POS], phi_hi[MAX_POS];    // floor(rho[i] * 2^64 / r)
};
struct RescaleConsts {
    RescaleParams p;
    Fr i_2S, i_pow_d, i_bound_d, i_pow_r, i_bound_r;   // canonical integers
    Fr m_2S, m_2SP, m_pow_d, m_bound_d, m_pow_r, m_bound_r;  // Montgomery forms
    LimbConsts lc;
};
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// one lane of the (converged) warp; ptxas then knows the guarded region runs single-threaded and
// feeds the uniform-operand bulk copies without a per-lane serialisation loop
__device__ __forceinline__ bool elect_one() {
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
    return is_leader != 0;
}

// Per-lane witness stream: values go to this lane's shared-memory row; every CH values (and at the
// end of the element) the warp ships its 32 rows to their places in out_wit.  cp.async.bulk takes
// uniform operands, so ONE elected lane issues the 32 row copies of the warp (per-lane issue would be
// serialised by the compiler into a 32-trip loop of ~14 instructions each -- measured: 44 % of all
// executed instructions); the warp's 32 elements are consecutive, so row r goes to gbase + r*W.
template <int CH, int NBUF>
struct WitnessStreamT {
    static constexpr int ROW_U4 = CH * 2 + 1;  // staging row in 16-byte units: CH witnesses + 16 B skew (conflict-free)
    // Staging layout: [buffer][lane][ROW_U4].  Consecutive lanes are ROW_U4 = 2*CH + 1 sixteen-byte units apart (4 banks
    // mod 32 for CH = 8), so the 8 lanes of a quarter-warp store to 8 disjoint bank groups: conflict-free.  (Keeping a
    // lane's NBUF rows adjacent instead doubles the lane stride to 8 banks mod 32 and makes every store 2-way
    // conflicting: measured 16.2 M conflicts per N=1024 launch, half of all shared-memory wavefronts.)
    uint4* row0;      // this lane's row in buffer 0
    uint4* warp_row0; // lane 0's row in buffer 0 (the elected lane walks all 32)
    uint32_t buf_stride;  // distance between the buffers in 16-byte units (rows per CTA * ROW_U4; unused when NBUF == 1)
    Fr* gwarp;        // out_wit position of lane 0's element, advanced by every flush
    int W, valid;     // distance (in witnesses) between the stripes of consecutive lanes -- W for consecutive elements;
                      // lanes of this warp that hold a real element
    int buf, fill;

    __device__ __forceinline__ void put(const Fr& v) {
        uint4* s = row0 + buf * buf_stride + 2 * fill;
        s[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        s[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        if (++fill == CH) flush();
    }
    __device__ __forceinline__ void flush() {  // warp-uniform: every lane has the same `fill`
        if (fill == 0) return;
        // generic-proxy writes of every lane -> visible to the async proxy, then the elected lane ships
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
            const uint32_t bytes = (uint32_t)(fill * sizeof(Fr));
            uint32_t src = smem_addr(warp_row0 + buf * buf_stride);
            Fr* dst = gwarp;
            for (int r = 0; r < valid; r++) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src),
                             "r"(bytes)
                             : "memory");
                src += ROW_U4 * sizeof(uint4);
                dst += W;
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // the rows written next were shipped NBUF-1 flushes ago: wait until the engine has read them
            if (NBUF == 2)
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        gwarp += fill;
        fill = 0;
        if (NBUF == 2) buf ^= 1;
    }
};

// Same interface, plain coalesced stores: the warp reads its 32 staged rows back 16 bytes per lane -- half a warp per
// 256-byte row -- and writes them with streaming 16-byte stores (STG.128), i.e. every store instruction covers two
// contiguous 256-byte pieces of two stripes.  No copy engine, no async proxy: the staging rows are free again as soon as
// the loads have returned, so one buffer per lane is enough (half the shared memory of the bulk-copy stream).
template <int CH>
struct WitnessStreamStg {
    static constexpr int ROW_U4 = CH * 2 + 1;  // staging row in 16-byte units: CH witnesses + 16 B skew (conflict-free)
    uint4* row0;       // this lane's staging row
    uint4* warp_row0;  // lane 0's staging row
    Fr* gwarp;         // out_wit position of lane 0's element, advanced by every flush
    int W, valid;      // stripe distance in witnesses; lanes of this warp that hold a real element
    int lane, fill;

    __device__ __forceinline__ void put(const Fr& v) {
        uint4* s = row0 + 2 * fill;
        s[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        s[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        if (++fill == CH) flush();
    }
    __device__ __forceinline__ void flush() {  // warp-uniform: every lane has the same `fill`
        if (fill == 0) return;
        __syncwarp();
        const int chunk = lane & 15, half = lane >> 4;
        if (chunk < 2 * fill) {
#pragma unroll 4
            for (int r = half; r < valid; r += 2) {
                const uint4 v = warp_row0[r * ROW_U4 + chunk];
                __stcs(reinterpret_cast<uint4*>(gwarp + (size_t)r * W) + chunk, v);
            }
        }
        __syncwarp();  // the rows are rewritten by the next burst
        gwarp += fill;
        fill = 0;
    }
};

// Same interface, TMA tensor stores whose pieces sit on 256-BYTE-ALIGNED addresses (tuning switch rescale_store = 1).
// Why: tools/store_pattern.cu (no arithmetic) -- 32 stripes x 256-byte pieces per burst reach 5.3 TB/s when half of the pieces
// straddle a 256-byte boundary (every other stripe of W = 60 witnesses does: 1920 bytes = 7.5 x 256) and 5.9-6.1 TB/s when none
// does; and the elected lane's 32 bulk copies per burst are a sixth of all executed instructions.  Here the EVEN and the ODD
// stripes of a warp are rows of two 3-D tensors [128 B][halves of a stripe][elements / 2] (a "half" = 128 bytes = 4 witnesses):
//   class 0 (even lanes, stripe 256-byte aligned):    chunk j = witnesses [8j, 8j + 8)       = halves 2j, 2j + 1
//   class 1 (odd lanes, stripe starts 128 bytes in):  head    = witnesses [0, 4)             = half 0
//                                                     chunk j = witnesses [4 + 8j, 12 + 8j)  = halves 2j + 1, 2j + 2
// so every chunk is one 256-byte-aligned piece per stripe, and ONE instruction (a box of 128 B x 2 halves x 16 rows) ships a
// class's chunk; the two classes complete their chunks 4 witnesses apart.  A stripe's last box reaches past its end
// (W % 8 == 4): the TMA unit clips it, as it clips the rows of a partial last warp.  (Stores with a NEGATIVE coordinate
// fault, hence the separate 128-byte head of class 1.)  Staging per warp: two buffers of 16 rows x 256 B per class in the
// 128-byte-swizzled layout the tensor maps declare (16-byte chunk c of 128-byte line L at c ^ (L & 7)) + the head tile.
// Buffer reuse: a class writes into a buffer again right after shipping the other one; the classes alternate, so the copy
// that last read the buffer was committed two events earlier (wait until at most 2 groups are pending).
struct WitnessStreamTma {
    static constexpr int WARP_BYTES = 2 * 2 * 4096 + 2048;     // [class][buffer] tiles + the head tile of class 1
    // usable when every even stripe is 256-byte aligned, stripes are an odd number of halves and both classes have rows
    static bool usable(const void* out_wit, int W, size_t count) {
        return (W & 7) == 4 && (reinterpret_cast<uintptr_t>(out_wit) & 255u) == 0 && count >= 2 && count < (1ull << 31);
    }
    uint8_t* wbase;                      // this warp's staging (1024-byte aligned)
    const CUtensorMap *map_even, *map_odd, *map_head;
    int row0;                            // first row of this warp in the two class tensors (= first element / 2)
    int lane;
    uint32_t cls, rho;                   // this lane's class and its row within the class
    uint32_t J0, J1;                     // chunks shipped so far per class (buffer parity), warp-uniform
    uint32_t done0, done1, count;        // per element: chunks shipped per class, witnesses written

    __device__ __forceinline__ void init(uint8_t* warp_stage, const CUtensorMap* me, const CUtensorMap* mo, const CUtensorMap* mh) {
        lane = threadIdx.x & 31;
        wbase = warp_stage;
        map_even = me;
        map_odd = mo;
        map_head = mh;
        cls = lane & 1;
        rho = lane >> 1;
        J0 = J1 = 0;
    }
    __device__ __forceinline__ void begin(int first_element) {
        row0 = first_element >> 1;
        done0 = done1 = 0;
        count = 0;
    }
    __device__ __forceinline__ void commit_and_wait() {
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
    __device__ __forceinline__ void put(const Fr& v) {
        // where witness `count` of this lane's element lives
        uint8_t* line;
        uint32_t sw, c;
        if (cls == 1u && count < 4u) {                                    // head tile: [16 rows][128 B]
            line = wbase + 4 * 4096 + rho * 128;
            sw = rho & 7u;
            c = 2u * count;
        } else {
            const uint32_t pos = cls ? (8u * J1 + count - 4u) & 15u : (8u * J0 + count) & 15u;
            const uint32_t buf = pos >> 3, s = pos & 7u, half = s >> 2;
            line = wbase + (cls * 2 + buf) * 4096 + rho * 256 + half * 128;
            sw = (2u * rho + half) & 7u;
            c = 2u * (s & 3u);
        }
        *reinterpret_cast<uint4*>(line + ((c ^ sw) << 4)) = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        *reinterpret_cast<uint4*>(line + (((c + 1u) ^ sw) << 4)) = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        count++;
        const uint32_t ph = count & 7u;
        if (ph == 0u || ph == 4u) {                                       // warp-uniform: a class has completed a chunk
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
                if (ph == 0u) {
                    const uint32_t src = smem_addr(wbase + (0 * 2 + ((J0 + done0) & 1u)) * 4096);
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_even),
                                 "r"(0), "r"(2 * (int)done0), "r"(row0), "r"(src)
                                 : "memory");
                } else if (count == 4u) {
                    const uint32_t src = smem_addr(wbase + 4 * 4096);
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_head),
                                 "r"(0), "r"(0), "r"(row0), "r"(src)
                                 : "memory");
                } else {
                    const uint32_t src = smem_addr(wbase + (1 * 2 + ((J1 + done1) & 1u)) * 4096);
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_odd),
                                 "r"(0), "r"(2 * (int)done1 + 1), "r"(row0), "r"(src)
                                 : "memory");
                }
                commit_and_wait();
            }
            __syncwarp();
            if (ph == 0u) done0++;
            else if (count != 4u) done1++;
        }
    }
    // end of the element (count = W, W % 8 == 4): class 0 has 4 witnesses left -- its box reaches one half past the stripe
    // (clipped); class 1's last chunk ended exactly here and went out in put()
    __device__ __forceinline__ void flush() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (elect_one()) {
            const uint32_t src = smem_addr(wbase + (0 * 2 + ((J0 + done0) & 1u)) * 4096);
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map_even), "r"(0),
                         "r"(2 * (int)done0), "r"(row0), "r"(src)
                         : "memory");
            commit_and_wait();
        }
        __syncwarp();
        J0 += done0 + 1u;
        J1 += done1;
    }
};

__device__ __forceinline__ Fr shr_small(const Fr& y, int s) {  // 1 <= s <= 32
    Fr o;
#pragma unroll
    for (int j = 0; j < 7; j++) o.l[j] = __funnelshift_rc(y.l[j], y.l[j + 1], s);
    o.l[7] = __funnelshift_rc(y.l[7], 0u, s);
    return o;
}

// RangeChip::range_check(x, n*lb): limbs l_i and running sums s_i = x mod 2^(lb*(i+1)), Montgomery form.
// RS_LIMB_PAIRS (experiment, off): two limbs per trip, so that four single-limb Montgomery steps are independent of each
// other and of the running sum.  Measured on the N = 1024 rescale: 471 us against 404 us one limb at a time -- the extra
// live values cost more than the instruction-level parallelism gains.
#ifndef RS_LIMB_PAIRS
#define RS_LIMB_PAIRS 0
#endif
// FAST: the running sums through fr::SmallSum (needs k.fast_sums; the kernels are instantiated both ways).
template <bool FAST = false, class WS>
__device__ __forceinline__ void stream_range_check(WS& ws, const LimbConsts& k, Fr y, int n) {
    if (n == 1) return;
    if constexpr (FAST) {
        fr::SmallSum ss;
        fr::small_sum_init(ss);
        {
            const uint32_t l = y.l[0] & k.lb_mask;
            y = shr_small(y, k.lb);
            fr::small_sum_add(ss, l, k.rho[0], k.phi_lo[0], k.phi_hi[0]);
            ws.put(fr::mont_mul_small(l, k.c[0]));
        }
        for (int i = 1; i < n; i++) {
            const uint32_t l = y.l[0] & k.lb_mask;
            y = shr_small(y, k.lb);
            ws.put(fr::mont_mul_small(l, k.c[0]));
            fr::small_sum_add(ss, l, k.rho[i], k.phi_lo[i], k.phi_hi[i]);
            ws.put(fr::small_sum_value(ss));
        }
        return;
    }
    Fr sum;
    {
        const uint32_t l = y.l[0] & k.lb_mask;
        y = shr_small(y, k.lb);
        sum = fr::mont_mul_small(l, k.c[0]);
        ws.put(sum);
    }
    int i = 1;
#if RS_LIMB_PAIRS
    for (; i + 1 < n; i += 2) {
        const uint32_t l0 = y.l[0] & k.lb_mask;
        y = shr_small(y, k.lb);
        const uint32_t l1 = y.l[0] & k.lb_mask;
        y = shr_small(y, k.lb);
        const Fr ml0 = fr::mont_mul_small(l0, k.c[0]);
        const Fr t0 = fr::mont_mul_small(l0, k.c[i]);
        const Fr ml1 = fr::mont_mul_small(l1, k.c[0]);
        const Fr t1 = fr::mont_mul_small(l1, k.c[i + 1]);
        const Fr s0 = fr::add_fast(sum, t0);
        sum = fr::add_fast(s0, t1);
        ws.put(ml0);
        ws.put(s0);
        ws.put(ml1);
        ws.put(sum);
    }
#endif
    for (; i < n; i++) {
        const uint32_t l = y.l[0] & k.lb_mask;
        y = shr_small(y, k.lb);
        const Fr ml = fr::mont_mul_small(l, k.c[0]);
        ws.put(ml);
        sum = fr::add_fast(sum, fr::mont_mul_small(l, k.c[i]));
        ws.put(sum);
    }
}

// RangeChip::check_big_less_than_safe(x, B)
template <bool FAST = false, class WS>
__device__ __forceinline__ void stream_cbls(WS& ws, const LimbConsts& k, const Fr& x_int,
                                            const Fr& x_mont, int n, const Fr& i_pow, const Fr& i_bound,
                                            const Fr& m_pow, const Fr& m_bound) {
    stream_range_check<FAST>(ws, k, x_int, n);
    const Fr chk_int = fr::sub_fast(fr::add_fast(x_int, i_pow), i_bound);  // x + 2^bits - B (mod r)
    const Fr m_xp = fr::add_fast(x_mont, m_pow);
    ws.put(fr::sub_fast(m_xp, m_bound));
    ws.put(m_xp);
    stream_range_check<FAST>(ws, k, chk_int, n);
}

// One element of rescale_matrix: the W witnesses of signed_div_scale(c) go to `ws` in assignment order, the quotient
// (Montgomery form) is returned.  Warp-uniform control flow (every lane streams the same number of witnesses).
template <bool FAST = false, class WS>
__device__ __forceinline__ Fr rescale_element(WS& ws, const RescaleConsts& k, const Fr& am) {
    const Fr a = fr::mont_reduce_fast(am);                  // canonical integer (reduction only: 80 instead of 132 IMAD.WIDE)
    const Fr ash = fr::add_fast(a, k.i_2S);                 // gate.add(a, Constant(2^S))
    const Fr div = fr::shr(ash, k.p.P);                     // div_mod_floor by 2^P
    const Fr rem = fr::low_bits(ash, k.p.P);
    const Fr m_div = fr::to_mont_fast(div);
    const Fr m_rem = fr::to_mont_fast(rem);
    ws.put(fr::add_fast(am, k.m_2S));
    ws.put(m_rem);
    ws.put(m_div);
    stream_cbls<FAST>(ws, k.lc, div, m_div, k.p.n_d, k.i_pow_d, k.i_bound_d, k.m_pow_d, k.m_bound_d);
    stream_cbls<FAST>(ws, k.lc, rem, m_rem, k.p.n_r, k.i_pow_r, k.i_bound_r, k.m_pow_r, k.m_bound_r);
    const Fr q = fr::sub_fast(m_div, k.m_2SP);              // gate.sub(div, Constant(2^(S-P)))
    ws.put(q);
    ws.flush();
    return q;
}

}  // namespace rs

// host side (rescale.cu): the constants of one (P, lb, S, A) configuration; returns W, or -1 if the parameters are out
// of range or need more limb positions than the staged kernels support
int make_rescale_consts(int P, int lb, int S, int A, rs::RescaleConsts* out);

}  // namespace h2svd
