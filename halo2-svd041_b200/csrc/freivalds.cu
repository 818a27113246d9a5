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
        if (e & 1u) acc = fr::mont_mul(acc, base);
        e >>= 1;
        if (e) base = fr::mont_mul(base, base);
    }
    st_fr_cs(out + i, acc);
}

// out[row*len + j] = sum_{t<=j} a[row*len + t] * v[row*v_row_stride + t]
__global__ void __launch_bounds__(256)
mat_vec_prefix_kernel(const Fr* __restrict__ a, const Fr* __restrict__ v, Fr* __restrict__ out, size_t rows,
                      size_t len, size_t v_row_stride) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t row = warp; row < rows; row += nwarps) {
        const Fr* ar = a + row * len;
        const Fr* vr = v + row * v_row_stride;
        Fr* orow = out + row * len;
        Fr carry = fr::zero();
        for (size_t base = 0; base < len; base += 32) {
            const size_t j = base + lane;
            Fr p = fr::zero();
            if (j < len) p = fr::mont_mul(ldg_fr(ar + j), ldg_fr(vr + j));
            // inclusive scan over the 32 lanes (Hillis-Steele, field adds)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                Fr t = shfl_up_fr(p, d);
                if (lane >= d) p = fr::add(p, t);
            }
            p = fr::add(p, carry);
            if (j < len) st_fr_cs(orow + j, p);
            carry = shfl_fr(p, 31);
        }
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
            acc = fr::mont_mul(acc, acc);
            if ((word >> bit) & 1u) acc = fr::mont_mul(acc, a);
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

int launch_mat_vec_prefix(h2svd_ctx* ctx, const Fr* a, const Fr* v, size_t rows, size_t len, size_t v_row_stride,
                          Fr* out) {
    if (rows == 0 || len == 0) return H2SVD_OK;
    const size_t warps_per_block = 8;
    size_t blocks = (rows + warps_per_block - 1) / warps_per_block;
    const size_t max_blocks = (size_t)ctx->sm_count * 8;  // grid-stride beyond that
    if (blocks > max_blocks) blocks = max_blocks;
    mat_vec_prefix_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, v, out, rows, len, v_row_stride);
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
