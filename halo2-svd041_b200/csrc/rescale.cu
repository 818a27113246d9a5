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
// All of it is shifts and masks on the canonical integer (the divisor is a power of two), followed
// by a conversion of every emitted value to Montgomery form.  One thread per element; each witness is
// written as one aligned 32-byte sector.
#include "common.cuh"

namespace h2svd {

namespace {

struct RescaleParams {
    int P, lb, S, A, n_d, n_r, W;
};

__device__ __forceinline__ Fr* emit(Fr* w, const Fr& x_int) {
    st_fr_cs(w, fr::to_mont(x_int));
    return w + 1;
}

// RangeChip::range_check(x, n*lb): limbs are the low n chunks of the canonical value
__device__ __forceinline__ Fr* emit_range_check(Fr* w, const Fr& x, int n, int lb) {
    if (n == 1) return w;
    for (int i = 0; i < n; i++) {
        w = emit(w, fr::low_bits(fr::shr(x, lb * i), lb));
        if (i >= 1) w = emit(w, fr::low_bits(x, lb * (i + 1)));
    }
    return w;
}

// RangeChip::check_big_less_than_safe(x, bound), n = ceil(bound.bits()/lb)
__device__ __forceinline__ Fr* emit_cbls(Fr* w, const Fr& x, const Fr& bound, int n, int lb) {
    w = emit_range_check(w, x, n, lb);
    const Fr xp = fr::add(x, fr::pow2(n * lb));  // x + 2^bits      (mod r)
    const Fr chk = fr::sub(xp, bound);           // x + 2^bits - B  (mod r)
    w = emit(w, chk);
    w = emit(w, xp);
    return emit_range_check(w, chk, n, lb);
}

__global__ void __launch_bounds__(128)
rescale_kernel(const Fr* __restrict__ cs, Fr* __restrict__ out_q, Fr* __restrict__ out_wit, size_t count,
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
    return 4 + 4 * (nd + nr);
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
    rescale_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(cs, out_q, out_wit, count, p);
    H2SVD_LAUNCH_CHECK(ctx);
    return H2SVD_OK;
}

}  // namespace h2svd
