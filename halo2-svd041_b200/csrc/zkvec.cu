// K5/K6 and input helpers: element-wise ZkVector witnesses, quantization, validation.
//   zkvec_sub      -- Witness cell of fpchip.qsub = gate.sub (reference src/matrix/mod.rs:143-146)
//   isqrt_fixed    -- value model of fpchip.qsqrt (:130, :163; SURVEY.md A.6, parity unpinned)
//   quantize       -- FixedPointChip041::quantization (:36, :245; SURVEY.md A.5, PDF Eq. 11)
//   check_canonical-- rejects limbs >= r at the boundary
// (The inner-product running sums of :79-100 reuse mat_vec_prefix_kernel in freivalds.cu with a
//  per-row second operand.)
#include "common.cuh"

namespace h2svd {

namespace {

__global__ void sub_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b, Fr* __restrict__ out, size_t count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        st_fr_cs(out + i, fr::sub(ldg_fr(a + i), ldg_fr(b + i)));
}

// floor(sqrt(x << P)) for x < 2^128 by the restoring (digit-by-digit) method on 256-bit integers
__global__ void isqrt_kernel(const Fr* __restrict__ a, Fr* __restrict__ out, size_t count, int P, int* flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Fr x = fr::from_mont(ldg_fr(a + i));
    if (x.l[4] | x.l[5] | x.l[6] | x.l[7]) {
        atomicExch(flag, 1);  // operand out of the modelled range
        st_fr(out + i, fr::zero());
        return;
    }
    // num = x << P  (P <= 63): limb t takes bits from source limbs t-w and t-w-1
    Fr num = fr::zero();
    {
        const int w = P >> 5, b = P & 31;
        uint32_t src[4] = {x.l[0], x.l[1], x.l[2], x.l[3]};
#pragma unroll
        for (int t = 0; t < 8; t++) {
            uint32_t lo = 0, hi = 0;  // lo = src[t-w-1], hi = src[t-w]
#pragma unroll
            for (int s = 0; s < 4; s++) {
                if (s == t - w - 1) lo = src[s];
                if (s == t - w) hi = src[s];
            }
            num.l[t] = b ? ((hi << b) | (lo >> (32 - b))) : hi;
        }
    }
    Fr res = fr::zero();
    for (int bitpos = 254; bitpos >= 0; bitpos -= 2) {
        const Fr bit = fr::pow2(bitpos);
        Fr t;
        fr::add_n<8>(t.l, res.l, bit.l);  // res + bit (< 2^256)
        res = fr::shr(res, 1);
        Fr d;
        const uint32_t borrow = fr::sub_n<8>(d.l, num.l, t.l);
        if (!borrow) {  // num >= res + bit
            num = d;
            fr::add_n<8>(res.l, res.l, bit.l);
        }
    }
    st_fr(out + i, fr::to_mont(res));
}

__global__ void quantize_kernel(const double* __restrict__ x, Fr* __restrict__ out, size_t count, int P,
                                int* flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double v = x[i];
        // round-half-away-from-zero of |x| * 2^P (PDF Eq. 11, Rust f64::round).  The scaling by a power of two is exact and
        // so is y - floor(y); forming y + 0.5 instead would tie-to-even for odd y in [2^52, 2^53) and round 0.49999999999999994 up.
        const double y = fabs(v) * scalbn(1.0, P);
        double mag = floor(y);
        if (y - mag >= 0.5) mag += 1.0;
        Fr q = fr::zero();
        if (!(mag < scalbn(1.0, 127))) {
            atomicExch(flag, 1);  // NaN / inf / out of range
        } else if (mag >= 1.0) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(mag);
            const int exp = (int)((bits >> 52) & 0x7ff) - 1075;  // mag = mant * 2^exp
            const unsigned long long mant = (bits & 0xfffffffffffffull) | (1ull << 52);
            // place the 53-bit mantissa at bit offset exp (exp may be negative: mag is an integer,
            // so the dropped bits are zero)
            if (exp <= 0) {
                const unsigned long long m = mant >> (-exp);
                q.l[0] = (uint32_t)m;
                q.l[1] = (uint32_t)(m >> 32);
            } else {
                const int w = exp >> 5, b = exp & 31;
                const uint32_t m0 = (uint32_t)mant, m1 = (uint32_t)(mant >> 32);
                const uint32_t p0 = m0 << b;
                const uint32_t p1 = b ? ((m1 << b) | (m0 >> (32 - b))) : m1;
                const uint32_t p2 = b ? (m1 >> (32 - b)) : 0u;
#pragma unroll
                for (int t = 0; t < 8; t++) q.l[t] = (t == w) ? p0 : (t == w + 1) ? p1 : (t == w + 2) ? p2 : 0u;
            }
            if (v < 0.0) {  // negatives are stored as r - q
                Fr m;
#pragma unroll
                for (int t = 0; t < 8; t++) m.l[t] = fr::modulus(t);
                Fr neg;
                fr::sub_n<8>(neg.l, m.l, q.l);
                q = neg;
            }
        }
        st_fr(out + i, fr::to_mont(q));
    }
}

__global__ void check_canonical_kernel(const Fr* __restrict__ x, size_t count, int* flag) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        bad |= !fr::is_canonical(ldg_fr(x + i));
    if (bad) atomicExch(flag, 1);
}

unsigned grid_for(h2svd_ctx* ctx, size_t count, int block) {
    size_t blocks = (count + block - 1) / block;
    const size_t cap = (size_t)ctx->sm_count * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

}  // namespace

int launch_sub(h2svd_ctx* ctx, const Fr* a, const Fr* b, size_t count, Fr* out) {
    if (count == 0) return H2SVD_OK;
    sub_kernel<<<grid_for(ctx, count, 256), 256, 0, ctx->stream>>>(a, b, out, count);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_isqrt(h2svd_ctx* ctx, const Fr* a, size_t count, int P, Fr* out) {
    if (P < 1 || P > 63) {
        set_error("isqrt_fixed: precision_bits out of range");
        return H2SVD_EINVAL;
    }
    if (count == 0) return H2SVD_OK;
    isqrt_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(a, out, count, P, ctx->d_flag);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_quantize(h2svd_ctx* ctx, const double* x, size_t count, int P, Fr* out) {
    if (P < 1 || P > 63) {
        set_error("quantize: precision_bits out of range");
        return H2SVD_EINVAL;
    }
    if (count == 0) return H2SVD_OK;
    quantize_kernel<<<grid_for(ctx, count, 256), 256, 0, ctx->stream>>>(x, out, count, P, ctx->d_flag);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

int launch_check_canonical(h2svd_ctx* ctx, const Fr* x, size_t count, int* d_flag) {
    if (count == 0) return H2SVD_OK;
    check_canonical_kernel<<<grid_for(ctx, count, 256), 256, 0, ctx->stream>>>(x, count, d_flag);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd
