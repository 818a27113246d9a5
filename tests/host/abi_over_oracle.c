/* TEST-ONLY stand-in for libh2svd_b200: implements the host-pointer entry points that
 * include/h2svd_zk.hpp calls on top of the C oracle, so the host mirror's cell bookkeeping can be
 * exercised on the CPU-only build box (pytest -m "not gpu").  Never built into or shipped with the
 * product; the GPU tests link the same mirror against the real CUDA library instead. */
#include <stdlib.h>
#include <string.h>

#include "../../include/h2svd_b200.h"
#include "../../oracle/fr_oracle.h"

struct h2svd_ctx { int dummy; };
static const char *g_err = "";
#define O(p) ((orc_fr *)(p))
#define OC(p) ((const orc_fr *)(p))
static int rc(int r, const char *what) { if (r) { g_err = what; return H2SVD_EINVAL; } return H2SVD_OK; }

const char *h2svd_last_error(void) { return g_err; }
int h2svd_create(h2svd_ctx **out, int device, void *stream) { (void)device; (void)stream; *out = calloc(1, sizeof(**out)); return H2SVD_OK; }
void h2svd_destroy(h2svd_ctx *c) { free(c); }
int h2svd_fr_matmul(h2svd_ctx *c, const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *o, size_t n, size_t k, size_t m, int bt) {
    (void)c; if (bt) return rc(1, "b_transposed unsupported in the test stand-in");
    return rc(orc_field_mat_mul(OC(a), OC(b), O(o), n, k, m, 0, n, 1), "field_mat_mul");
}
int h2svd_freivalds_witness(h2svd_ctx *c, const h2svd_fr *a, const h2svd_fr *b, const h2svd_fr *cs, const h2svd_fr *g, size_t n,
                            size_t k, size_t m, h2svd_fr *pw, h2svd_fr *pcv, h2svd_fr *pbv, h2svd_fr *pabv, h2svd_fr *d,
                            h2svd_fr *z, h2svd_fr *inv) {
    (void)c;
    return rc(orc_freivalds_witness(OC(a), OC(b), OC(cs), OC(g), n, k, m, O(pw), O(pcv), O(pbv), O(pabv), O(d), O(z), O(inv), 1), "freivalds");
}
int h2svd_mat_vec_prefix(h2svd_ctx *c, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t len, h2svd_fr *out) {
    (void)c; return rc(orc_mat_vec_prefix(OC(a), OC(v), rows, len, O(out), 1), "mat_vec_prefix");
}
int h2svd_rescale_witness_count(int P, int lb, int S, int A) { return orc_rescale_witness_count(P, lb, S, A); }
int h2svd_rescale_witness(h2svd_ctx *c, const h2svd_fr *cs, size_t count, int P, int lb, int S, int A, h2svd_fr *q, h2svd_fr *w) {
    (void)c; return rc(orc_rescale_witness(OC(cs), count, P, lb, S, A, O(q), NULL, O(w), 1), "rescale");
}
int h2svd_zkvec_inner_prefix(h2svd_ctx *c, const h2svd_fr *x, const h2svd_fr *s, size_t batch, size_t len, h2svd_fr *out) {
    (void)c; return rc(orc_zkvec_inner_prefix(OC(x), OC(s), batch, len, O(out), 1), "zkvec_inner_prefix");
}
int h2svd_zkvec_sub(h2svd_ctx *c, const h2svd_fr *s, const h2svd_fr *x, size_t count, h2svd_fr *out) {
    (void)c; return rc(orc_zkvec_sub(OC(s), OC(x), count, O(out)), "zkvec_sub");
}
int h2svd_isqrt_fixed(h2svd_ctx *c, const h2svd_fr *a, size_t count, int P, h2svd_fr *out) {
    (void)c; return rc(orc_isqrt_fixed(OC(a), count, P, O(out)), "isqrt");
}
int h2svd_quantize(h2svd_ctx *c, const double *x, size_t count, int P, h2svd_fr *out) {
    (void)c; return rc(orc_quantize(x, count, P, O(out)), "quantize");
}
int h2svd_abs_less_than_witness_count(const uint64_t bnd[4], int lb, int with_diff) { return orc_abs_less_than_witness_count(bnd, lb, with_diff); }
int h2svd_abs_less_than_witness(h2svd_ctx *c, const h2svd_fr *x, const h2svd_fr *y, size_t count, const uint64_t bnd[4], int lb, h2svd_fr *w) {
    (void)c; return rc(orc_abs_less_than_witness(OC(x), y ? OC(y) : NULL, count, bnd, lb, O(w)), "abs_less_than");
}
int h2svd_range_check_witness_count(int bits, int lb) { return orc_range_check_witness_count(bits, lb); }
int h2svd_range_check_witness(h2svd_ctx *c, const h2svd_fr *x, size_t count, int bits, int lb, h2svd_fr *w) {
    (void)c; return rc(orc_range_check_witness(OC(x), count, bits, lb, O(w)), "range_check");
}
int h2svd_mat_times_diag(h2svd_ctx *c, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t lda, size_t cols_v, h2svd_fr *out) {
    (void)c; return rc(orc_mat_times_diag(OC(a), OC(v), rows, lda, cols_v, O(out)), "mat_times_diag");
}
int h2svd_host_fr_from_canonical(const uint64_t x[4], h2svd_fr *out) { orc_fr_from_canonical(O(out), x); return H2SVD_OK; }
void h2svd_host_fr_to_canonical(const h2svd_fr *a, uint64_t out[4]) { orc_fr_to_canonical(out, OC(a)); }
void h2svd_host_fr_add(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *o) { orc_fr_add(O(o), OC(a), OC(b)); }
void h2svd_host_fr_sub(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *o) { orc_fr_sub(O(o), OC(a), OC(b)); }
void h2svd_host_fr_mul(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *o) { orc_fr_mul(O(o), OC(a), OC(b)); }
