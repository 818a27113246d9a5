/* C ORACLE for the ZkMatrix / ZkVector witness path -- TEST INFRASTRUCTURE ONLY.
 *
 * A CPU restatement of the reference's algorithm (reference = /root/reference,
 * Rust, not buildable here: no cargo/rustc, un-vendored crates).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline.
 * The product (halo2-svd041_b200/) never links or calls it.
 *
 * PARITY STATUS: parity unpinned at the third-party boundary (see
 * oracle/pyoracle.py header and DESIGN.md).  This file is cross-checked
 * bit-for-bit against oracle/pyoracle.py (Python big ints) in tests/.
 *
 * All field elements: halo2curves bn256::Fr wire format -- 4 x u64
 * little-endian limbs of x * 2^256 mod r, canonical (< r).
 */
#ifndef FR_ORACLE_H
#define FR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } orc_fr;

/* scalar field ops (Montgomery domain in, Montgomery domain out) */
void orc_fr_mul(orc_fr *o, const orc_fr *a, const orc_fr *b);
void orc_fr_add(orc_fr *o, const orc_fr *a, const orc_fr *b);
void orc_fr_sub(orc_fr *o, const orc_fr *a, const orc_fr *b);
void orc_fr_from_canonical(orc_fr *o, const uint64_t x[4]); /* integer < r -> Montgomery */
void orc_fr_to_canonical(uint64_t x[4], const orc_fr *a);   /* Montgomery -> integer */
int orc_fr_is_canonical(const orc_fr *a);

/* reference src/matrix/mod.rs:510-537 field_mat_mul, i-j-k order.
 * a: n x k, b: k x m, c: n x m, row-major contiguous.  Computes rows
 * [row0, row1) of c only (so a caller can time a bounded sample); threads > 1
 * splits those rows over OpenMP threads (the reference itself is 1 thread). */
int orc_field_mat_mul(const orc_fr *a, const orc_fr *b, orc_fr *c, size_t n, size_t k, size_t m,
                      size_t row0, size_t row1, int threads);

/* reference src/matrix/mod.rs:316-326: out[i] = gamma^i, i < d (sequential chain) */
int orc_gamma_powers(const orc_fr *gamma, size_t d, orc_fr *out);

/* reference src/matrix/mod.rs:574-599 + GateChip::inner_product running sums:
 * out[i*len + j] = sum_{t<=j} a[i*len+t] * v[t] */
int orc_mat_vec_prefix(const orc_fr *a, const orc_fr *v, size_t rows, size_t len, orc_fr *out,
                       int threads);

/* Witness-count helper for signed_div_scale: W = 4 + f(n_d) + f(n_r), f(n) = 4n (n >= 2) or 2 (n == 1) */
int orc_rescale_witness_count(int precision_bits, int lookup_bits, int shift_bits, int a_num_bits);

/* reference src/matrix/mod.rs:354-375 rescale_matrix -> per element
 * FixedPointChip041::signed_div_scale.  out_q[count],
 * out_rem[count] (may be NULL), out_wit[count * W] in cell order. */
int orc_rescale_witness(const orc_fr *cs, size_t count, int precision_bits, int lookup_bits,
                        int shift_bits, int a_num_bits, orc_fr *out_q, orc_fr *out_rem,
                        orc_fr *out_wit, int threads);

/* reference src/matrix/mod.rs:79-100 ZkVector::inner_product running sums for
 * `batch` independent (x, self) pairs of length len: out[b*len+j] =
 * sum_{t<=j} x[b][t] * self[b][t] */
int orc_zkvec_inner_prefix(const orc_fr *x, const orc_fr *self, size_t batch, size_t len,
                           orc_fr *out, int threads);

/* reference src/matrix/mod.rs:143-146: diff[i] = self[i] - x[i] (qsub = gate.sub) */
int orc_zkvec_sub(const orc_fr *self, const orc_fr *x, size_t count, orc_fr *out);

/* reference src/matrix/mod.rs:425-459 check_abs_less_than / check_mat_diff witnesses:
 * [x - y]? , t = d + (bnd - 1), check_big_less_than_safe(t, 2*bnd - 1); bnd = canonical integer */
int orc_abs_less_than_witness_count(const uint64_t bnd[4], int lookup_bits, int with_diff);
int orc_abs_less_than_witness(const orc_fr *x, const orc_fr *y, size_t count, const uint64_t bnd[4],
                              int lookup_bits, orc_fr *out_wit);
/* RangeChip::range_check(x, range_bits) witnesses (reference :185-216 callers) */
int orc_range_check_witness_count(int range_bits, int lookup_bits);
int orc_range_check_witness(const orc_fr *x, size_t count, int range_bits, int lookup_bits,
                            orc_fr *out_wit);
/* reference :610-627 mat_times_diag_mat values */
int orc_mat_times_diag(const orc_fr *a, const orc_fr *v, size_t rows, size_t lda, size_t cols_v,
                       orc_fr *out);

/* FixedPointChip041::quantization: sign-magnitude round-half-up */
int orc_quantize(const double *x, size_t count, int precision_bits, orc_fr *out);

/* qsqrt model (unpinned): out = floor(sqrt(a * 2^P)), a < 2^128 */
int orc_isqrt_fixed(const orc_fr *a, size_t count, int precision_bits, orc_fr *out);

/* Full Freivalds witness (reference src/matrix/mod.rs:299-342): every
 * Witness-kind cell of verify_mul in assignment order, split per array:
 *   powers[m]; prefix_cv[n*m]; prefix_bv[k*m]; prefix_abv[n*k];
 *   diff[n] (= cs_v - ab_v), is_zero[n], inv[n]                         */
int orc_freivalds_witness(const orc_fr *a, const orc_fr *b, const orc_fr *cs, const orc_fr *gamma,
                          size_t n, size_t k, size_t m, orc_fr *powers, orc_fr *prefix_cv,
                          orc_fr *prefix_bv, orc_fr *prefix_abv, orc_fr *diff, orc_fr *is_zero,
                          orc_fr *inv, int threads);

#ifdef __cplusplus
}
#endif
#endif
