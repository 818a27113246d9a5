// K2/K3: Freivalds witnesses of ZkMatrix::verify_mul (reference src/matrix/mod.rs:299-342) and the
// ZkVector::inner_product running sums (:79-100).
//
//  * gamma powers (:316-326): the reference chains v_i = v_{i-1} * gamma sequentially; here thread i
//    computes gamma^i by square-and-multiply (field arithmetic is exact, so the values are identical).
//  * mat-vec with EVERY running sum emitted (field_mat_vec_mul :574-599 -> GateChip::inner_product):
//    one warp per row; lane l takes elements l, l+32, ... so loads/stores are fully coalesced
//    (32 lanes x 32 B = 1 KiB contiguous per step), products are reduced per lane and the running sum
//    is a warp-shuffle inclusive scan in the field plus the carry from the previous 32-block.
//  * is_equal cells (:339-341).
#include "common.cuh"
#include "fr_acc.cuh"
#include "fr_fast.cuh"

namespace h2svd {

namespace {

__device__ __forceinline__ Fr shfl_up_fr(const Fr& v, int delta) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], delta);
    return r;
}
__device__ __forceinline__ Fr shfl_fr(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
    return r;
}

__global__ void gamma_powers_kernel(const Fr* __restrict__ gamma, Fr* __restrict__ out, unsigned d) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    Fr base = ldg_fr(gamma);
    Fr acc = fr::one();
    unsigned e = i;
    while (e) {
        if (e & 1u) acc = fr::mont_mul_fast(acc, base);
        e >>= 1;
        if (e) base = fr::mont_mul_fast(base, base);
    }
    st_fr_cs(out + i, acc);
}

// inclusive scan of one field element per lane over the first `width` lanes (width = 32 or 8)
template <int WIDTH>
__device__ __forceinline__ Fr warp_scan_fr(Fr p, int lane) {
#pragma unroll
    for (int d = 1; d < WIDTH; d <<= 1) {
        const Fr t = shfl_up_fr(p, d);
        const Fr s = fr::add_fast(p, t);
        if (lane >= d) p = s;
    }
    return p;
}

constexpr int MV_WARPS = 8;   // warps per CTA
constexpr int MV_C = 4;       // 32-element chunks per warp per tile
constexpr int MV_MAX_JOBS = 2;

struct MvJob {
    const Fr* a;     // rows x len
    Fr* out;         // rows x len running sums
    Fr* totals;      // rows (last running sum of every row) or nullptr
    size_t rows;
};
struct MvJobs {
    MvJob job[MV_MAX_JOBS];
    int njobs;
};

// out[row*len + j] = sum_{t<=j} a[row*len + t] * v[row*v_row_stride + t]   for every row of every job
// (all jobs share v, len and v_row_stride).  WPR warps cooperate on one row: inside a tile of
// WPR*128 elements warp w owns the contiguous elements [w*128, (w+1)*128) as four coalesced 32-lane
// chunks; it keeps its four un-offset running sums in registers, the warps exchange their segment
// totals through shared memory, and every running sum is written exactly once.
template <int WPR>
__device__ __forceinline__ void mat_vec_prefix_body(const MvJobs& jobs, const Fr* __restrict__ v, size_t len, size_t v_row_stride) {
    constexpr int RPC = MV_WARPS / WPR;       // rows per CTA pass
    constexpr int TILE = WPR * 32 * MV_C;
    __shared__ Fr wtot[MV_WARPS];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int rloc = warp / WPR, wr = warp % WPR;
    size_t total_rows = 0;
#pragma unroll
    for (int q = 0; q < MV_MAX_JOBS; q++) total_rows += q < jobs.njobs ? jobs.job[q].rows : 0;

    for (size_t row_base = (size_t)blockIdx.x * RPC; row_base < total_rows; row_base += (size_t)gridDim.x * RPC) {
        size_t row = row_base + rloc;
        const bool active = row < total_rows;
        const Fr* a = nullptr;
        Fr* out = nullptr;
        Fr* totals = nullptr;
        if (active) {
            int q = 0;
            if (jobs.njobs > 1 && row >= jobs.job[0].rows) {
                row -= jobs.job[0].rows;
                q = 1;
            }
            a = jobs.job[q].a + row * len;
            out = jobs.job[q].out + row * len;
            totals = jobs.job[q].totals ? jobs.job[q].totals + row : nullptr;
        }
        const Fr* vr = v + (active ? row * v_row_stride : 0);
        Fr carry = fr::zero();  // sum of all complete tiles of this row (identical in every warp of the row)
        for (size_t t0 = 0; t0 < len; t0 += TILE) {
            const size_t seg0 = t0 + (size_t)wr * 32 * MV_C;
            Fr res[MV_C];
            Fr run = fr::zero();
#pragma unroll
            for (int c = 0; c < MV_C; c++) {
                const size_t j = seg0 + c * 32 + lane;
                Fr p = fr::zero();
                if (seg0 + c * 32 < len) {  // warp-uniform
                    if (active && j < len) p = fr::mont_mul_fast(ldg_fr(a + j), ldg_fr(vr + j));
                    p = warp_scan_fr<32>(p, lane);
                    p = fr::add_fast(p, run);
                    run = shfl_fr(p, 31);
                }
                res[c] = p;
            }
            if (WPR > 1) {
                if (lane == 0) st_fr(&wtot[warp], run);
                __syncthreads();
                // lanes 0..WPR-1 scan the segment totals of this row
                Fr t = fr::zero();
                if (lane < WPR) t = ld_fr(&wtot[rloc * WPR + lane]);
                t = warp_scan_fr<WPR>(t, lane);
                const Fr incl_prev = shfl_fr(t, wr > 0 ? wr - 1 : 0);
                const Fr tile_total = shfl_fr(t, WPR - 1);
                Fr off = carry;
                if (wr > 0) off = fr::add_fast(off, incl_prev);
                carry = fr::add_fast(carry, tile_total);
                __syncthreads();  // wtot is rewritten by the next tile
#pragma unroll
                for (int c = 0; c < MV_C; c++) {
                    const size_t j = seg0 + c * 32 + lane;
                    if (active && j < len) st_fr_cs(out + j, fr::add_fast(res[c], off));
                }
            } else {
#pragma unroll
                for (int c = 0; c < MV_C; c++) {
                    const size_t j = seg0 + c * 32 + lane;
                    if (active && j < len) st_fr_cs(out + j, fr::add_fast(res[c], carry));
                }
                carry = fr::add_fast(carry, run);
            }
        }
        if (totals && wr == 0 && lane == 0) st_fr(totals, carry);
    }
}
template <int WPR>
__global__ void __launch_bounds__(MV_WARPS * 32, 2)
mat_vec_prefix_kernel(const __grid_constant__ MvJobs jobs, const Fr* __restrict__ v, size_t len, size_t v_row_stride) {
    mat_vec_prefix_body<WPR>(jobs, v, len, v_row_stride);
}
// The same kernel capped at 96 registers and ~1 KB of shared memory: one CTA of it fits on an SM NEXT TO the three resident
// CTAs of the rescale kernel (104 registers x 384 threads, 209 KB of staging), whose store-bound warps leave half the issue
// slots idle (experiment behind the tuning switch "step_schedule").
template <int WPR>
__global__ void __maxnreg__(96)
mat_vec_prefix_lowreg_kernel(const __grid_constant__ MvJobs jobs, const Fr* __restrict__ v, size_t len, size_t v_row_stride) {
    mat_vec_prefix_body<WPR>(jobs, v, len, v_row_stride);
}

// ---------------------------------------------------------------------------------------------------
// Consecutive-element version of the same computation for rows of >= 128 elements (the BASELINE shapes):
// ONE WARP PER ROW, the row walked in 128-element tiles, each lane owning FOUR CONSECUTIVE elements of the tile.
// The strided mapping of the kernel above needs a 5-level warp-shuffle scan per element (5 field adds + 40
// shuffles); with consecutive elements the running sum is one field add per element, one shuffle scan covers
// 128 elements, and a second, coalesced pass adds each lane's offset: ~300 instead of ~580 instructions per
// element, leaving the 132 IMAD.WIDE of the Montgomery product (which every canonical running sum needs) as
// the bound.  Warps never synchronise with each other (no __syncthreads, no flags): the 24 resident warps of an
// SM sit in different phases (loading / multiplying / scanning / storing), which keeps the multiplier pipe fed.
//  * each warp stages its a / v tiles with cp.async (16-byte LDGSTS, coalesced on the global side) into its
//    own 128-byte-XOR-swizzled buffers (16-byte chunk q lives at q ^ ((q >> 3) & 7)): the lane-consecutive
//    reads of pass 1, the stores of the un-offset sums and the coalesced reads of pass 2 are conflict-free.
//  * the un-offset running sums overwrite the a tile; pass 2 adds the owner lane's offset and stores straight
//    to global in element order.
constexpr int MT_WARPS = 4;                      // warps per CTA (a container only: warps are independent)
constexpr int MT_EPL = 4;                        // consecutive elements per lane
constexpr int MT_SEG = 32 * MT_EPL;              // 128 elements (4 KB) per tile
constexpr int MT_SEG_U4 = MT_SEG * 2;            // 16-byte chunks per tile buffer
constexpr int MT_WARP_U4 = 2 * MT_SEG_U4 + 64;   // a tile + v tile + 32 lane offsets
constexpr size_t MT_SMEM = (size_t)MT_WARPS * MT_WARP_U4 * sizeof(uint4);   // 36 KB -> 6 CTAs = 24 warps per SM

__device__ __forceinline__ int mt_swz(int q) { return q ^ ((q >> 3) & 7); }
__device__ __forceinline__ Fr mt_ld(const uint4* buf, int e) {
    const uint4 lo = buf[mt_swz(2 * e)], hi = buf[mt_swz(2 * e + 1)];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
__device__ __forceinline__ void mt_st(uint4* buf, int e, const Fr& v) {
    buf[mt_swz(2 * e)] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    buf[mt_swz(2 * e + 1)] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// one warp stages `count` (<= MT_SEG) elements of src into a swizzled tile buffer
__device__ __forceinline__ void mt_stage(uint4* buf, const Fr* src, int count, int lane) {
    const uint4* g = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int r = 0; r < 2 * MT_EPL; r++) {
        const int q = lane + 32 * r;
        if (q < 2 * count) {
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(buf + mt_swz(q)));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + q) : "memory");
        }
    }
}

// pass 1 of a tile: products and un-offset running sums of this lane's MT_EPL consecutive elements (written over the a
// tile); returns the lane total.  X2: the products are taken two at a time with interleaved carry chains.
template <bool X2>
__device__ __forceinline__ Fr mt_pass1(uint4* sa, const uint4* sv, int lane, int nseg) {
    Fr run = fr::zero();
    if (X2) {
#pragma unroll
        for (int i = 0; i < MT_EPL; i += 2) {
            const int e = MT_EPL * lane + i;
            if (e + 1 < nseg) {
                Fr p0, p1;
                fr::mont_mul_fast_x2(mt_ld(sa, e), mt_ld(sv, e), mt_ld(sa, e + 1), mt_ld(sv, e + 1), p0, p1);
                run = fr::add_fast(run, p0);
                mt_st(sa, e, run);
                run = fr::add_fast(run, p1);
                mt_st(sa, e + 1, run);
            } else if (e < nseg) {
                run = fr::add_fast(run, fr::mont_mul_fast(mt_ld(sa, e), mt_ld(sv, e)));
                mt_st(sa, e, run);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < MT_EPL; i++) {
            const int e = MT_EPL * lane + i;
            if (e < nseg) {
                run = fr::add_fast(run, fr::mont_mul_fast(mt_ld(sa, e), mt_ld(sv, e)));
                mt_st(sa, e, run);
            }
        }
    }
    return run;
}

template <bool X2>
__global__ void __launch_bounds__(MT_WARPS * 32)
mat_vec_prefix_tile_kernel(const __grid_constant__ MvJobs jobs, const Fr* __restrict__ v, size_t len,
                           size_t v_row_stride) {
    extern __shared__ __align__(128) uint4 mt_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint4* sa = mt_smem + (size_t)w * MT_WARP_U4;
    uint4* sv = sa + MT_SEG_U4;
    Fr* soff = reinterpret_cast<Fr*>(sv + MT_SEG_U4);
    size_t total_rows = 0;
#pragma unroll
    for (int q = 0; q < MV_MAX_JOBS; q++) total_rows += q < jobs.njobs ? jobs.job[q].rows : 0;
    const size_t warps_total = (size_t)gridDim.x * MT_WARPS;

    for (size_t grow = (size_t)blockIdx.x * MT_WARPS + w; grow < total_rows; grow += warps_total) {
        size_t row = grow;
        int jq = 0;
        if (jobs.njobs > 1 && row >= jobs.job[0].rows) {
            row -= jobs.job[0].rows;
            jq = 1;
        }
        const Fr* a = jobs.job[jq].a + row * len;
        Fr* out = jobs.job[jq].out + row * len;
        const Fr* vr = v + row * v_row_stride;
        Fr carry = fr::zero();  // sum of the complete tiles of this row
        for (size_t t0 = 0; t0 < len; t0 += MT_SEG) {
            const int nseg = (int)(len - t0 < (size_t)MT_SEG ? len - t0 : (size_t)MT_SEG);
            mt_stage(sa, a + t0, nseg, lane);
            mt_stage(sv, vr + t0, nseg, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            // pass 1: products and un-offset running sums of this lane's four consecutive elements
            const Fr run = mt_pass1<X2>(sa, sv, lane, nseg);
            const Fr incl = warp_scan_fr<32>(run, lane);
            Fr off = shfl_up_fr(incl, 1);  // exclusive over the lanes
            if (lane == 0) off = fr::zero();
            off = fr::add_fast(off, carry);
            carry = fr::add_fast(carry, shfl_fr(incl, 31));
            st_fr(&soff[lane], off);
            __syncwarp();
            // pass 2: coalesced -- element e gets the offset of its owner lane e / MT_EPL
#pragma unroll
            for (int r = 0; r < MT_EPL; r++) {
                const int e = lane + 32 * r;
                if (e < nseg) st_fr_cs(out + t0 + e, fr::add_fast(mt_ld(sa, e), ld_fr(&soff[e / MT_EPL])));
            }
            __syncwarp();  // the buffers are restaged by the next tile
        }
        if (lane == 0 && jobs.job[jq].totals) st_fr(jobs.job[jq].totals + row, carry);
    }
}

// ---------------------------------------------------------------------------------------------------
// Few-rows version of the tile kernel (the Freivalds mat-vecs of one N = 1024 product, or a 128-row slab of it: 1024 /
// 128 rows leave one-warp-per-row with 7 / 1 warps per SM -- measured 10.7 % warps active, 33 % issue active).  A CTA of
// SEGS warps takes ONE row: warp w owns the 128-element tiles w, w + SEGS, ... of it and computes their products and
// tile-local running sums exactly like the tile kernel; the SEGS tile totals of a round meet in shared memory (one
// __syncthreads per round, double-buffered by round parity), every warp adds the totals of the tiles before its own to
// its lane offsets, and the second, coalesced pass stores.  Same arithmetic per element, SEGS times the parallelism.
template <int SEGS, bool X2>
__global__ void __launch_bounds__(SEGS * 32)
mat_vec_prefix_seg_kernel(const __grid_constant__ MvJobs jobs, const Fr* __restrict__ v, size_t len, size_t v_row_stride) {
    extern __shared__ __align__(128) uint4 mt_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint4* sa = mt_smem + (size_t)w * MT_WARP_U4;
    uint4* sv = sa + MT_SEG_U4;
    Fr* soff = reinterpret_cast<Fr*>(sv + MT_SEG_U4);
    Fr* stot = reinterpret_cast<Fr*>(mt_smem + (size_t)SEGS * MT_WARP_U4);   // [2][SEGS] tile totals
    size_t total_rows = 0;
#pragma unroll
    for (int q = 0; q < MV_MAX_JOBS; q++) total_rows += q < jobs.njobs ? jobs.job[q].rows : 0;
    const size_t rounds = (len + (size_t)SEGS * MT_SEG - 1) / ((size_t)SEGS * MT_SEG);

    for (size_t grow = blockIdx.x; grow < total_rows; grow += gridDim.x) {   // CTA-uniform: barriers inside are safe
        size_t row = grow;
        int jq = 0;
        if (jobs.njobs > 1 && row >= jobs.job[0].rows) {
            row -= jobs.job[0].rows;
            jq = 1;
        }
        const Fr* a = jobs.job[jq].a + row * len;
        Fr* out = jobs.job[jq].out + row * len;
        const Fr* vr = v + row * v_row_stride;
        Fr carry = fr::zero();  // sum of all tiles of the previous rounds (identical in every warp)
        for (size_t rd = 0; rd < rounds; rd++) {
            const size_t t0 = (rd * SEGS + w) * (size_t)MT_SEG;
            const int nseg = t0 >= len ? 0 : (int)(len - t0 < (size_t)MT_SEG ? len - t0 : (size_t)MT_SEG);   // warp-uniform
            Fr incl = fr::zero(), off = fr::zero();
            if (nseg > 0) {
                mt_stage(sa, a + t0, nseg, lane);
                mt_stage(sv, vr + t0, nseg, lane);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                const Fr run = mt_pass1<X2>(sa, sv, lane, nseg);
                incl = warp_scan_fr<32>(run, lane);
                off = shfl_up_fr(incl, 1);  // exclusive over the lanes
                if (lane == 0) off = fr::zero();
            }
            Fr* tot = stot + (rd & 1) * SEGS;
            if (lane == 31) st_fr(&tot[w], incl);   // tile total (zero for a tile past the end of the row)
            __syncthreads();
            // totals of the tiles before this warp's own, and of the whole round
            Fr before = carry;
#pragma unroll
            for (int q = 0; q < SEGS; q++) {
                const Fr t = ld_fr(&tot[q]);
                if (q < w) before = fr::add_fast(before, t);
                carry = fr::add_fast(carry, t);
            }
            if (nseg > 0) {
                st_fr(&soff[lane], fr::add_fast(off, before));
                __syncwarp();
#pragma unroll
                for (int r = 0; r < MT_EPL; r++) {
                    const int e = lane + 32 * r;
                    if (e < nseg) st_fr_cs(out + t0 + e, fr::add_fast(mt_ld(sa, e), ld_fr(&soff[e / MT_EPL])));
                }
                __syncwarp();  // the buffers are restaged by the next round
            }
        }
        if (threadIdx.x == 0 && jobs.job[jq].totals) st_fr(jobs.job[jq].totals + row, carry);
    }
}

// Row totals only (no running sums): totals[row] = sum_t a[row][t] * v[t].  Used by row-sharded callers for (B v): every
// rank needs all k totals as the second operand of A.(Bv) (reference src/matrix/mod.rs:337) but emits the running-sum
// witnesses of its own rows of B only -- computing the totals redundantly is cheaper than a collective.  No canonical
// intermediate is needed, so the products are accumulated lazily (fr_acc.cuh: 64 IMAD.WIDE per element, one reduction
// per lane) -- the same value as the last running sum of mat_vec_prefix, since the canonical representative is unique.
constexpr int MVT_WPR = 4;      // warps per row: 1024 rows of one N = 1024 operand would otherwise leave 7 warps per SM
constexpr int MVT_ROWS = 2;     // rows per CTA
__global__ void __launch_bounds__(MVT_WPR * MVT_ROWS * 32)
mat_vec_totals_kernel(const Fr* __restrict__ a, const Fr* __restrict__ v, Fr* __restrict__ totals, size_t rows, size_t len) {
    __shared__ Fr part[MVT_ROWS][MVT_WPR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = warp / MVT_WPR, wr = warp % MVT_WPR;
    for (size_t row0 = (size_t)blockIdx.x * MVT_ROWS; row0 < rows; row0 += (size_t)gridDim.x * MVT_ROWS) {   // CTA-uniform
        const size_t row = row0 + rl;
        fr::WideAcc w;
        fr::acc_clear(w);
        if (row < rows) {
            const Fr* ar = a + row * len;
            for (size_t j = (size_t)wr * 32 + lane; j < len; j += MVT_WPR * 32) {
                const Fr x = ldg_fr(ar + j), y = ldg_fr(v + j);
                fr::mul_acc(w, x.l, y.l);
            }
        }
        Fr p = fr::acc_finalize(w);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            Fr t;
#pragma unroll
            for (int i = 0; i < 8; i++) t.l[i] = __shfl_down_sync(0xffffffffu, p.l[i], d);
            p = fr::add_fast(p, t);
        }
        if (lane == 0) st_fr(&part[rl][wr], p);
        __syncthreads();
        if (wr == 0 && lane == 0 && row < rows) {
            Fr s = ld_fr(&part[rl][0]);
#pragma unroll
            for (int q = 1; q < MVT_WPR; q++) s = fr::add_fast(s, ld_fr(&part[rl][q]));
            st_fr(totals + row, s);
        }
        __syncthreads();   // part[] is rewritten by the next pair of rows
    }
}

__global__ void gather_kernel(const Fr* __restrict__ src, Fr* __restrict__ out, size_t count, size_t stride,
                              size_t offset) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) st_fr(out + i, ldg_fr(src + i * stride + offset));
}

// a^(r-2) by square-and-multiply (only ever reached for a dishonest c_s: diff != 0)
__device__ Fr fr_inverse(const Fr& a) {
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = fr::modulus(i);
    e[0] -= 2u;  // low limb is 0xf0000001: no borrow
    Fr acc = fr::one();
    for (int w = 7; w >= 0; w--) {
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) word = (i == w) ? e[i] : word;
        for (int bit = 31; bit >= 0; bit--) {
            acc = fr::mont_mul_fast(acc, acc);
            if ((word >> bit) & 1u) acc = fr::mont_mul_fast(acc, a);
        }
    }
    return acc;
}

__global__ void is_equal_kernel(const Fr* __restrict__ x, const Fr* __restrict__ y, Fr* __restrict__ diff,
                                Fr* __restrict__ is_zero, Fr* __restrict__ inv, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr d = fr::sub(ldg_fr(x + i), ldg_fr(y + i));  // gate.sub -> Witness(a - b)
    st_fr(diff + i, d);
    if (fr::is_zero(d)) {
        st_fr(is_zero + i, fr::one());
        st_fr(inv + i, fr::one());
    } else {
        st_fr(is_zero + i, fr::zero());
        st_fr(inv + i, fr_inverse(d));
    }
}

}  // namespace

int launch_gamma_powers(h2svd_ctx* ctx, const Fr* gamma, size_t d, Fr* out) {
    if (d == 0) return H2SVD_OK;
    if (d > 0xffffffffull) {
        set_error("gamma_powers: d too large");
        return H2SVD_EINVAL;
    }
    gamma_powers_kernel<<<(unsigned)((d + 127) / 128), 128, 0, ctx->stream>>>(gamma, out, (unsigned)d);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

template <int WPR>
static int launch_mv(h2svd_ctx* ctx, const MvJobs& jobs, size_t total_rows, const Fr* v, size_t len, size_t vs) {
    constexpr int RPC = MV_WARPS / WPR;
    size_t blocks = (total_rows + RPC - 1) / RPC;
    const size_t cap = (size_t)ctx->sm_count * 2 * 4;  // 2 resident CTAs per SM, grid-stride beyond 4 waves
    if (blocks > cap) blocks = cap;
    if (ctx->tune.matvec_coreside)
        mat_vec_prefix_lowreg_kernel<WPR><<<(unsigned)blocks, MV_WARPS * 32, 0, ctx->stream>>>(jobs, v, len, vs);
    else
        mat_vec_prefix_kernel<WPR><<<(unsigned)blocks, MV_WARPS * 32, 0, ctx->stream>>>(jobs, v, len, vs);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

template <int SEGS, bool X2>
static int launch_mv_seg_x(h2svd_ctx* ctx, const MvJobs& jobs, size_t total_rows, const Fr* v, size_t len, size_t vs) {
    constexpr size_t smem = (size_t)SEGS * MT_WARP_U4 * sizeof(uint4) + 2 * SEGS * sizeof(Fr);
    H2SVD_SET_SMEM(ctx, (mat_vec_prefix_seg_kernel<SEGS, X2>), smem);
    size_t blocks = total_rows;
    const size_t cap = (size_t)ctx->sm_count * 16;   // grid-stride beyond a few waves of resident CTAs
    if (blocks > cap) blocks = cap;
    mat_vec_prefix_seg_kernel<SEGS, X2><<<(unsigned)blocks, SEGS * 32, smem, ctx->stream>>>(jobs, v, len, vs);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}
template <int SEGS>
static int launch_mv_seg(h2svd_ctx* ctx, const MvJobs& jobs, size_t total_rows, const Fr* v, size_t len, size_t vs) {
    return ctx->tune.matvec_x2 ? launch_mv_seg_x<SEGS, true>(ctx, jobs, total_rows, v, len, vs)
                               : launch_mv_seg_x<SEGS, false>(ctx, jobs, total_rows, v, len, vs);
}

static int launch_mv_jobs(h2svd_ctx* ctx, const MvJobs& jobs, const Fr* v, size_t len, size_t vs) {
    size_t total_rows = 0;
    for (int q = 0; q < jobs.njobs; q++) total_rows += jobs.job[q].rows;
    if (total_rows == 0 || len == 0) return H2SVD_OK;
    // few long rows: several warps per row (mat_vec_prefix_seg_kernel); "matvec_seg": -1 auto, 0 never, 1 whenever len >= 256
    const bool few_rows = total_rows < (size_t)ctx->sm_count * 16;
    if (len >= 256 && ctx->tune.matvec_warp == 0 && !ctx->tune.matvec_coreside && (ctx->tune.matvec_seg > 0 || (ctx->tune.matvec_seg < 0 && few_rows))) {
        const int segs = ctx->tune.matvec_segs > 0 ? ctx->tune.matvec_segs : (len >= 1024 ? 8 : len >= 512 ? 4 : 2);
        if (segs >= 8) return launch_mv_seg<8>(ctx, jobs, total_rows, v, len, vs);
        if (segs >= 4) return launch_mv_seg<4>(ctx, jobs, total_rows, v, len, vs);
        return launch_mv_seg<2>(ctx, jobs, total_rows, v, len, vs);
    }
    if (len >= 128 && total_rows >= (size_t)ctx->sm_count * 4 && ctx->tune.matvec_warp == 0 && !ctx->tune.matvec_coreside) {
        // one warp per row, consecutive elements per lane (mat_vec_prefix_tile_kernel); needs enough rows to
        // give every SM a few warps, otherwise the row-splitting kernel below is the better fit
        size_t blocks = (total_rows + MT_WARPS - 1) / MT_WARPS;
        const size_t cap = (size_t)ctx->sm_count * 5;  // 5 CTAs (96 regs) of 4 independent warps per SM, grid-stride beyond
        if (blocks > cap) blocks = cap;
        if (ctx->tune.matvec_x2) {
            H2SVD_SET_SMEM(ctx, mat_vec_prefix_tile_kernel<true>, MT_SMEM);
            mat_vec_prefix_tile_kernel<true><<<(unsigned)blocks, MT_WARPS * 32, MT_SMEM, ctx->stream>>>(jobs, v, len, vs);
        } else {
            H2SVD_SET_SMEM(ctx, mat_vec_prefix_tile_kernel<false>, MT_SMEM);
            mat_vec_prefix_tile_kernel<false><<<(unsigned)blocks, MT_WARPS * 32, MT_SMEM, ctx->stream>>>(jobs, v, len, vs);
        }
        H2SVD_LAUNCH_CHECK(ctx);
        return H2SVD_OK;
    }
    // warps per row: enough to cover the row with 128-element segments, and enough to fill the GPU
    int wpr = 1;
    while (wpr < MV_WARPS && (size_t)wpr * 128 < len) wpr <<= 1;
    switch (wpr) {
        case 1: return launch_mv<1>(ctx, jobs, total_rows, v, len, vs);
        case 2: return launch_mv<2>(ctx, jobs, total_rows, v, len, vs);
        case 4: return launch_mv<4>(ctx, jobs, total_rows, v, len, vs);
        default: return launch_mv<8>(ctx, jobs, total_rows, v, len, vs);
    }
}

int launch_mat_vec_prefix(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t len, size_t v_row_stride,
                          Fr* out, Fr* totals) {
    MvJobs jobs{};
    jobs.njobs = 1;
    jobs.job[0] = MvJob{a, out, totals, rows};
    return launch_mv_jobs(ctx, jobs, v, len, v_row_stride);
}

// two matrices against the same vector in one launch (C.v and B.v of verify_mul)
int launch_mat_vec_prefix2(h2svd_ctx* ctx, const Fr* a0, size_t rows0, Fr* out0, Fr* totals0, const Fr* a1,
                           size_t rows1, Fr* out1, Fr* totals1, const Fr* v, size_t len) {
    MvJobs jobs{};
    jobs.njobs = 2;
    jobs.job[0] = MvJob{a0, out0, totals0, rows0};
    jobs.job[1] = MvJob{a1, out1, totals1, rows1};
    return launch_mv_jobs(ctx, jobs, v, len, 0);
}

int launch_mat_vec_totals(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t len, Fr* totals) {
    if (rows == 0) return H2SVD_OK;
    size_t blocks = (rows + MVT_ROWS - 1) / MVT_ROWS;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    mat_vec_totals_kernel<<<(unsigned)blocks, MVT_WPR * MVT_ROWS * 32, 0, ctx->stream>>>(a, v, totals, rows, len);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_gather(h2svd_ctx* ctx, const Fr* src, size_t count, size_t stride, size_t offset, Fr* out) {
    if (count == 0) return H2SVD_OK;
    gather_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(src, out, count, stride, offset);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_is_equal(h2svd_ctx* ctx, const Fr* x, const Fr* y, size_t count, Fr* diff, Fr* is_zero, Fr* inv) {
    if (count == 0) return H2SVD_OK;
    is_equal_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(x, y, diff, is_zero, inv, count);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd

