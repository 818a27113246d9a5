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
__global__ void __launch_bounds__(MV_WARPS * 32, 2)
mat_vec_prefix_kernel(const __grid_constant__ MvJobs jobs, const Fr* __restrict__ v, size_t len, size_t v_row_stride) {
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
    mat_vec_prefix_kernel<WPR><<<(unsigned)blocks, MV_WARPS * 32, 0, ctx->stream>>>(jobs, v, len, vs);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

static int launch_mv_jobs(h2svd_ctx* ctx, const MvJobs& jobs, const Fr* v, size_t len, size_t vs) {
    size_t total_rows = 0;
    for (int q = 0; q < jobs.njobs; q++) total_rows += jobs.job[q].rows;
    if (total_rows == 0 || len == 0) return H2SVD_OK;
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
