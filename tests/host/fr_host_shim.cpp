// Host build of the product's field-arithmetic headers (fr.cuh / fr_acc.cuh) so that the exact
// code the GPU runs outside the PTX chains is unit-tested on the CPU against the oracle.
// Compiled by tests/test_fr_host.py with g++; never shipped.
#include <cstddef>
#include <cstring>

#include "../../halo2-svd041_b200/csrc/fr_acc.cuh"
#include "../../halo2-svd041_b200/csrc/fr_fast.cuh"
#include "../../halo2-svd041_b200/csrc/fr_kara.cuh"
#include "../../halo2-svd041_b200/csrc/tc_small.cuh"

using fr::Fr;

extern "C" {
void hs_mont_mul(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::mont_mul(a[i], b[i]); }
void hs_add(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::add(a[i], b[i]); }
void hs_sub(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::sub(a[i], b[i]); }
void hs_to_mont(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::to_mont(a[i]); }
void hs_from_mont(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::from_mont(a[i]); }
void hs_shr(const Fr* a, int s, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::shr(a[i], s); }
void hs_low_bits(const Fr* a, int bits, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::low_bits(a[i], bits); }
void hs_pow2(int s, Fr* o) { *o = fr::pow2(s); }
int hs_is_canonical(const Fr* a) { return fr::is_canonical(*a) ? 1 : 0; }
void hs_one(Fr* o) { *o = fr::one(); }
// carry-chain primitives (host restatement of the PTX blocks in fr_fast.cuh)
void hs_mont_mul_fast(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::mont_mul_fast(a[i], b[i]); }
void hs_mont_mul_small(const uint32_t* l, const Fr* c, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::mont_mul_small(l[i], c[i]); }
void hs_add_fast(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::add_fast(a[i], b[i]); }
void hs_sub_fast(const Fr* a, const Fr* b, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::sub_fast(a[i], b[i]); }
void hs_to_mont_fast(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::to_mont_fast(a[i]); }
void hs_from_mont_fast(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::from_mont_fast(a[i]); }
void hs_mont_mul_fast_x2(const Fr* a, const Fr* b, Fr* o, size_t n) {   // pairs (2i, 2i+1)
    for (size_t i = 0; i + 1 < n; i += 2) fr::mont_mul_fast_x2(a[i], b[i], a[i + 1], b[i + 1], o[i], o[i + 1]);
}
// running sums of n small multiples l[j] * rho[j] (fr::SmallSum): out[i] = sum_{j<=i} l[j] * rho[j] mod r
void hs_small_sums(const uint32_t* l, const Fr* rho, const uint64_t* phi, Fr* out, size_t n) {
    fr::SmallSum s;
    fr::small_sum_init(s);
    for (size_t i = 0; i < n; i++) {
        fr::small_sum_add(s, l[i], rho[i], (uint32_t)phi[i], (uint32_t)(phi[i] >> 32));
        out[i] = fr::small_sum_value(s);
    }
}
void hs_mont_reduce_fast(const Fr* a, Fr* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fr::mont_reduce_fast(a[i]); }
// Small-operand tensor-core engine, arithmetic only (tc_small.cuh): balanced digits of every operand, the 17 signed
// diagonal sums a tile of s8 x s8 MMAs would leave in TMEM, signed carry, one Montgomery encode.  Returns 0 when an
// operand is out of the small range (|x| >= 2^70), else 1.
int hs_small_dot(const Fr* a, const Fr* b, size_t k, Fr* o) {
    int32_t D[2 * fr::SMALL_DIGITS - 1] = {0};
    for (size_t i = 0; i < k; i++) {
        uint32_t ta[3], tb[3];
        if (!fr::small_biased(a[i], ta) || !fr::small_biased(b[i], tb)) return 0;
        for (int p = 0; p < fr::SMALL_DIGITS; p++)
            for (int q = 0; q < fr::SMALL_DIGITS; q++)
                D[p + q] += (int32_t)(int8_t)fr::small_digit_bits(ta, p) * (int32_t)(int8_t)fr::small_digit_bits(tb, q);
    }
    uint32_t T[6];
    fr::carry_signed<2 * fr::SMALL_DIGITS - 1>(reinterpret_cast<const uint32_t*>(D), T);
    *o = fr::signed6_to_mont(T);
    return 1;
}
// the cheap small-operand test must agree with the full-reduction one, bit for bit (returns -1 on any difference)
int hs_small_fast_vs_full(const Fr* a, size_t n) {
    int nsmall = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t t0[3] = {0, 0, 0}, t1[3] = {0, 0, 0};
        const bool ok0 = fr::small_biased(a[i], t0), ok1 = fr::small_biased_fast(a[i], t1);
        if (ok0 != ok1) return -1;
        if (ok0 && (t0[0] != t1[0] || t0[1] != t1[1] || t0[2] != t1[2])) return -1;
        nsmall += ok0 ? 1 : 0;
    }
    return nsmall;
}
int hs_small_ok(const Fr* a) { uint32_t t[3]; return fr::small_biased(*a, t) ? 1 : 0; }
// Karatsuba lazy dot product exactly as the Karatsuba mat-mul accumulates it (fr_kara.cuh)
void hs_kara_dot(const Fr* a, const Fr* b, size_t k, Fr* o) {
    fr::KAcc p0, p1, p2;
    fr::kacc_clear(p0); fr::kacc_clear(p1); fr::kacc_clear(p2);
    for (size_t i = 0; i < k; i++) {
        const fr::KOp x = fr::ksplit(a[i]), y = fr::ksplit(b[i]);
        fr::kmul_acc(p0, x.lo, y.lo);
        fr::kmul_acc(p2, x.hi, y.hi);
        fr::kmul_acc(p1, x.s, y.s);
    }
    *o = fr::kara_finalize(p0, p1, p2);
}
void hs_kara_repeat(const Fr* a, const Fr* b, size_t k, Fr* o) {
    fr::KAcc p0, p1, p2;
    fr::kacc_clear(p0); fr::kacc_clear(p1); fr::kacc_clear(p2);
    const fr::KOp x = fr::ksplit(*a), y = fr::ksplit(*b);
    for (size_t i = 0; i < k; i++) {
        fr::kmul_acc(p0, x.lo, y.lo);
        fr::kmul_acc(p2, x.hi, y.hi);
        fr::kmul_acc(p1, x.s, y.s);
    }
    *o = fr::kara_finalize(p0, p1, p2);
}
// lazy dot product exactly as the mat-mul inner loop accumulates it (host fallback of chain4)
void hs_lazy_dot(const Fr* a, const Fr* b, size_t k, Fr* o) {
    fr::WideAcc w;
    fr::acc_clear(w);
    for (size_t i = 0; i < k; i++) fr::mul_acc(w, a[i].l, b[i].l);
    *o = fr::acc_finalize(w);
}
// C = A*B with the same lazy accumulation (row-major), for shape/ragged checks of the arithmetic
void hs_lazy_matmul(const Fr* A, const Fr* B, Fr* C, size_t n, size_t k, size_t m) {
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < m; j++) {
            fr::WideAcc w;
            fr::acc_clear(w);
            for (size_t t = 0; t < k; t++) fr::mul_acc(w, A[i * k + t].l, B[t * m + j].l);
            C[i * m + j] = fr::acc_finalize(w);
        }
}
// worst-case accumulator stress: k copies of the same product
void hs_lazy_repeat(const Fr* a, const Fr* b, size_t k, Fr* o) {
    fr::WideAcc w;
    fr::acc_clear(w);
    for (size_t i = 0; i < k; i++) fr::mul_acc(w, a->l, b->l);
    *o = fr::acc_finalize(w);
}
}
