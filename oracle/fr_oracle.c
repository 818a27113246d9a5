/* C ORACLE -- TEST INFRASTRUCTURE ONLY (see fr_oracle.h).
 *
 * Scalar BN254-Fr arithmetic the way halo2curves does it on a CPU (4 x u64
 * Montgomery, 4x4 schoolbook + word-by-word reduction, one conditional
 * subtract), and the reference's loops restated on top of it.  Each function
 * cites the reference file:line it follows (paths relative to /root/reference).
 */
#include "fr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;

static const uint64_t MODULUS[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                    0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const orc_fr MONT_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                                 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const orc_fr MONT_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL,
                                0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const uint64_t INV = 0xc2e1f593efffffffULL; /* -r^-1 mod 2^64 */

/* ---- 256-bit integer helpers ---------------------------------------- */
static int u256_geq(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static uint64_t u256_add(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        o[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static uint64_t u256_sub(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - borrow;
        o[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static void u256_shr(uint64_t o[4], const uint64_t a[4], unsigned s) {
    uint64_t t[4] = {0, 0, 0, 0};
    unsigned w = s / 64, b = s % 64;
    for (unsigned i = 0; i + w < 4; i++) {
        t[i] = a[i + w] >> b;
        if (b && i + w + 1 < 4) t[i] |= a[i + w + 1] << (64 - b);
    }
    memcpy(o, t, sizeof t);
}
static void u256_pow2(uint64_t o[4], unsigned s) {
    memset(o, 0, 32);
    o[s / 64] = 1ULL << (s % 64);
}
static void u256_low_bits(uint64_t o[4], const uint64_t a[4], unsigned bits) {
    for (unsigned i = 0; i < 4; i++) {
        if (bits >= 64 * (i + 1)) o[i] = a[i];
        else if (bits <= 64 * i) o[i] = 0;
        else o[i] = a[i] & ((1ULL << (bits - 64 * i)) - 1);
    }
}
/* (a + b) mod r for a, b < r */
static void int_add_mod(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    u256_add(o, a, b); /* < 2r < 2^255: no carry out */
    if (u256_geq(o, MODULUS)) u256_sub(o, o, MODULUS);
}
/* (a - b) mod r for a, b < r */
static void int_sub_mod(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) {
    if (u256_sub(o, a, b)) u256_add(o, o, MODULUS);
}

/* ---- Montgomery field ops ------------------------------------------- */
void orc_fr_add(orc_fr *o, const orc_fr *a, const orc_fr *b) { int_add_mod(o->l, a->l, b->l); }
void orc_fr_sub(orc_fr *o, const orc_fr *a, const orc_fr *b) { int_sub_mod(o->l, a->l, b->l); }

/* mac: (lo, carry) = x + y*z + carry -- the halo2curves building block */
#define MAC(lo, x, y, z, carry)                          \
    do {                                                 \
        u128 _t = (u128)(y) * (z) + (x) + (carry);       \
        (lo) = (uint64_t)_t;                             \
        (carry) = (uint64_t)(_t >> 64);                  \
    } while (0)

void orc_fr_mul(orc_fr *o, const orc_fr *a, const orc_fr *b) {
    const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3];
    const uint64_t b0 = b->l[0], b1 = b->l[1], b2 = b->l[2], b3 = b->l[3];
    uint64_t r0, r1, r2, r3, r4, r5, r6, r7, c;
    /* 4x4 schoolbook */
    c = 0; MAC(r0, 0, a0, b0, c); MAC(r1, 0, a0, b1, c); MAC(r2, 0, a0, b2, c); MAC(r3, 0, a0, b3, c); r4 = c;
    c = 0; MAC(r1, r1, a1, b0, c); MAC(r2, r2, a1, b1, c); MAC(r3, r3, a1, b2, c); MAC(r4, r4, a1, b3, c); r5 = c;
    c = 0; MAC(r2, r2, a2, b0, c); MAC(r3, r3, a2, b1, c); MAC(r4, r4, a2, b2, c); MAC(r5, r5, a2, b3, c); r6 = c;
    c = 0; MAC(r3, r3, a3, b0, c); MAC(r4, r4, a3, b1, c); MAC(r5, r5, a3, b2, c); MAC(r6, r6, a3, b3, c); r7 = c;
    /* word-by-word Montgomery reduction */
    uint64_t k, junk, c2;
    k = r0 * INV; c = 0;
    MAC(junk, r0, k, MODULUS[0], c); MAC(r1, r1, k, MODULUS[1], c); MAC(r2, r2, k, MODULUS[2], c); MAC(r3, r3, k, MODULUS[3], c);
    { u128 t = (u128)r4 + c; r4 = (uint64_t)t; c2 = (uint64_t)(t >> 64); }
    k = r1 * INV; c = 0;
    MAC(junk, r1, k, MODULUS[0], c); MAC(r2, r2, k, MODULUS[1], c); MAC(r3, r3, k, MODULUS[2], c); MAC(r4, r4, k, MODULUS[3], c);
    { u128 t = (u128)r5 + c + c2; r5 = (uint64_t)t; c2 = (uint64_t)(t >> 64); }
    k = r2 * INV; c = 0;
    MAC(junk, r2, k, MODULUS[0], c); MAC(r3, r3, k, MODULUS[1], c); MAC(r4, r4, k, MODULUS[2], c); MAC(r5, r5, k, MODULUS[3], c);
    { u128 t = (u128)r6 + c + c2; r6 = (uint64_t)t; c2 = (uint64_t)(t >> 64); }
    k = r3 * INV; c = 0;
    MAC(junk, r3, k, MODULUS[0], c); MAC(r4, r4, k, MODULUS[1], c); MAC(r5, r5, k, MODULUS[2], c); MAC(r6, r6, k, MODULUS[3], c);
    { u128 t = (u128)r7 + c + c2; r7 = (uint64_t)t; c2 = (uint64_t)(t >> 64); }
    (void)junk;
    uint64_t res[4] = {r4, r5, r6, r7};
    if (c2 || u256_geq(res, MODULUS)) u256_sub(res, res, MODULUS);
    memcpy(o->l, res, 32);
}

void orc_fr_from_canonical(orc_fr *o, const uint64_t x[4]) {
    orc_fr t;
    memcpy(t.l, x, 32);
    orc_fr_mul(o, &t, &MONT_R2);
}
void orc_fr_to_canonical(uint64_t x[4], const orc_fr *a) {
    orc_fr one = {{1, 0, 0, 0}}, t;
    orc_fr_mul(&t, a, &one);
    memcpy(x, t.l, 32);
}
int orc_fr_is_canonical(const orc_fr *a) { return !u256_geq(a->l, MODULUS); }

static void fr_pow(orc_fr *o, const orc_fr *a, const uint64_t e[4]) {
    orc_fr acc = MONT_ONE;
    for (int i = 255; i >= 0; i--) {
        orc_fr_mul(&acc, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) orc_fr_mul(&acc, &acc, a);
    }
    *o = acc;
}
static void fr_inv(orc_fr *o, const orc_fr *a) {
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    u256_sub(e, MODULUS, two);
    fr_pow(o, a, e);
}
static int fr_is_zero(const orc_fr *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

/* ---- tiny pthread parallel-for (the reference is single-threaded; threads > 1
 * only serves the "all host cores" CPU-baseline leg) ------------------------ */
typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; size_t lo, hi; } par_job;
static void *par_tramp(void *p) {
    par_job *j = (par_job *)p;
    j->fn(j->lo, j->hi, j->ctx);
    return NULL;
}
static void par_for(size_t lo, size_t hi, int threads, range_fn fn, void *ctx) {
    if (threads <= 0) {
        long nc = sysconf(_SC_NPROCESSORS_ONLN);
        threads = nc > 0 ? (int)nc : 1;
    }
    size_t total = hi > lo ? hi - lo : 0;
    if ((size_t)threads > total) threads = (int)(total ? total : 1);
    if (threads <= 1) {
        fn(lo, hi, ctx);
        return;
    }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    par_job *jobs = (par_job *)malloc(sizeof(par_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn;
        jobs[t].ctx = ctx;
        jobs[t].lo = lo + total * (size_t)t / (size_t)threads;
        jobs[t].hi = lo + total * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&tid[t], NULL, par_tramp, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
}

/* ---- reference loops --------------------------------------------------- */

/* reference src/matrix/mod.rs:525-535 */
typedef struct { const orc_fr *a, *b; orc_fr *c; size_t k, m; } mm_ctx;
static void mm_rows(size_t lo, size_t hi, void *p) {
    mm_ctx *x = (mm_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        for (size_t j = 0; j < x->m; j++) {
            orc_fr elem = {{0, 0, 0, 0}};
            for (size_t t = 0; t < x->k; t++) {
                orc_fr pr;
                orc_fr_mul(&pr, &x->a[i * x->k + t], &x->b[t * x->m + j]); /* :530 */
                orc_fr_add(&elem, &elem, &pr);
            }
            x->c[i * x->m + j] = elem;
        }
    }
}
int orc_field_mat_mul(const orc_fr *a, const orc_fr *b, orc_fr *c, size_t n, size_t k, size_t m,
                      size_t row0, size_t row1, int threads) {
    if (row1 > n || row0 > row1) return -1;
    mm_ctx x = {a, b, c, k, m};
    par_for(row0, row1, threads, mm_rows, &x);
    return 0;
}

/* reference src/matrix/mod.rs:316-326 */
int orc_gamma_powers(const orc_fr *gamma, size_t d, orc_fr *out) {
    if (d == 0) return 0;
    out[0] = MONT_ONE; /* :318 */
    for (size_t i = 1; i < d; i++) orc_fr_mul(&out[i], &out[i - 1], gamma); /* :324 */
    return 0;
}

/* reference src/matrix/mod.rs:582-596; running sums per halo2-base inner_product */
typedef struct { const orc_fr *a, *v; orc_fr *out; size_t len; int v_per_row; } mv_ctx;
static void mv_rows(size_t lo, size_t hi, void *p) {
    mv_ctx *x = (mv_ctx *)p;
    for (size_t i = lo; i < hi; i++) {
        orc_fr s = {{0, 0, 0, 0}};
        const orc_fr *v = x->v_per_row ? x->v + i * x->len : x->v;
        for (size_t j = 0; j < x->len; j++) {
            orc_fr pr;
            orc_fr_mul(&pr, &x->a[i * x->len + j], &v[j]);
            orc_fr_add(&s, &s, &pr);
            x->out[i * x->len + j] = s;
        }
    }
}
int orc_mat_vec_prefix(const orc_fr *a, const orc_fr *v, size_t rows, size_t len, orc_fr *out,
                       int threads) {
    mv_ctx x = {a, v, out, len, 0};
    par_for(0, rows, threads, mv_rows, &x);
    return 0;
}

/* reference src/matrix/mod.rs:88-100 (operand order u = x, v = self) */
int orc_zkvec_inner_prefix(const orc_fr *x, const orc_fr *self, size_t batch, size_t len,
                           orc_fr *out, int threads) {
    mv_ctx c = {x, self, out, len, 1};
    par_for(0, batch, threads, mv_rows, &c);
    return 0;
}

/* reference src/matrix/mod.rs:144-146 */
int orc_zkvec_sub(const orc_fr *self, const orc_fr *x, size_t count, orc_fr *out) {
    for (size_t i = 0; i < count; i++) orc_fr_sub(&out[i], &self[i], &x[i]);
    return 0;
}

/* ---- signed_div_scale witness ------------------------ */
static int ceil_div(int a, int b) { return (a + b - 1) / b; }

int orc_rescale_witness_count(int P, int lb, int S, int A) {
    if (S < 0) S = 3 * P;
    if (A < 0) A = 4 * P;
    if (P < 1 || P > 63 || lb < 1 || lb > 32 || S < P || S > 252 || A <= P) return -1;
    int n_d = ceil_div(A - P + 1, lb), n_r = ceil_div(P + 1, lb);
    if (n_d * lb > 253 || n_r * lb > 253) return -1;
    /* per check_big_less_than_safe: range_check (2n-1) + chk, xp + range_check (2n-1) = 4n; only chk, xp when n == 1 */
    return 4 + (n_d >= 2 ? 4 * n_d : 2) + (n_r >= 2 ? 4 * n_r : 2);
}

static orc_fr *emit(orc_fr *w, const uint64_t x[4]) {
    orc_fr_from_canonical(w, x);
    return w + 1;
}
/* RangeChip::range_check(x, n*lb) witnesses: l0, l1, s1, l2, s2, ... */
static orc_fr *emit_range_check(orc_fr *w, const uint64_t x[4], int n, int lb) {
    if (n == 1) return w;
    uint64_t t[4], limb[4], s[4];
    for (int i = 0; i < n; i++) {
        u256_shr(t, x, (unsigned)(lb * i));
        u256_low_bits(limb, t, (unsigned)lb);
        w = emit(w, limb);
        if (i >= 1) {
            u256_low_bits(s, x, (unsigned)(lb * (i + 1)));
            w = emit(w, s);
        }
    }
    return w;
}
/* RangeChip::check_big_less_than_safe(x, bound) with bound.bits() = bbits */
static orc_fr *emit_cbls(orc_fr *w, const uint64_t x[4], const uint64_t bound[4], int bbits,
                         int lb) {
    int n = ceil_div(bbits, lb), bits = n * lb;
    w = emit_range_check(w, x, n, lb);
    uint64_t p2[4], xp[4], chk[4];
    u256_pow2(p2, (unsigned)bits);
    int_add_mod(xp, x, p2);       /* a + 2^bits           */
    int_sub_mod(chk, xp, bound);  /* a + 2^bits - b       */
    w = emit(w, chk);
    w = emit(w, xp);
    return emit_range_check(w, chk, n, lb);
}

typedef struct {
    const orc_fr *cs;
    orc_fr *out_q, *out_rem, *out_wit;
    int P, lb, S, A, W;
    uint64_t p2s[4], bound_d[4], bound_r[4], qoff[4];
} rs_ctx;
static void rs_elems(size_t lo, size_t hi, void *p) {
    rs_ctx *x = (rs_ctx *)p;
    for (size_t e = lo; e < hi; e++) {
        uint64_t a[4], ash[4], div[4], rem[4], q[4];
        orc_fr_to_canonical(a, &x->cs[e]);
        int_add_mod(ash, a, x->p2s);              /* gate.add(a, 2^S)                 */
        u256_shr(div, ash, (unsigned)x->P);       /* div_mod_floor by 2^P             */
        u256_low_bits(rem, ash, (unsigned)x->P);
        orc_fr *w = x->out_wit + e * (size_t)x->W;
        w = emit(w, ash);
        w = emit(w, rem);
        w = emit(w, div);
        w = emit_cbls(w, div, x->bound_d, x->A - x->P + 1, x->lb);
        w = emit_cbls(w, rem, x->bound_r, x->P + 1, x->lb);
        int_sub_mod(q, div, x->qoff);             /* gate.sub(div, 2^(S-P))           */
        w = emit(w, q);
        orc_fr_from_canonical(&x->out_q[e], q);
        if (x->out_rem) orc_fr_from_canonical(&x->out_rem[e], rem);
    }
}
int orc_rescale_witness(const orc_fr *cs, size_t count, int P, int lb, int S, int A,
                        orc_fr *out_q, orc_fr *out_rem, orc_fr *out_wit, int threads) {
    if (S < 0) S = 3 * P;
    if (A < 0) A = 4 * P;
    int W = orc_rescale_witness_count(P, lb, S, A);
    if (W < 0) return -1;
    rs_ctx x = {cs, out_q, out_rem, out_wit, P, lb, S, A, W, {0}, {0}, {0}, {0}};
    uint64_t one[4] = {1, 0, 0, 0};
    u256_pow2(x.p2s, (unsigned)S);
    u256_pow2(x.bound_d, (unsigned)(A - P));
    u256_add(x.bound_d, x.bound_d, one); /* 2^A / 2^P + 1 */
    u256_pow2(x.bound_r, (unsigned)P);
    u256_pow2(x.qoff, (unsigned)(S - P));
    par_for(0, count, threads, rs_elems, &x);
    return 0;
}

/* ---- range-check witnesses of the SVD verifier's helpers (reference src/matrix/mod.rs:425-501, :185-216, :610-627) ---- */
static int u256_bits(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--)
        if (a[i]) return 64 * i + (64 - __builtin_clzll(a[i]));
    return 0;
}
int orc_abs_less_than_witness_count(const uint64_t bnd[4], int lb, int with_diff) {
    uint64_t two[4], bound[4], one[4] = {1, 0, 0, 0};
    if (lb < 1 || lb > 32 || u256_bits(bnd) == 0 || u256_bits(bnd) > 250) return -1;
    u256_add(two, bnd, bnd);
    u256_sub(bound, two, one);
    int n = ceil_div(u256_bits(bound), lb);
    if (n * lb > 253) return -1;
    return (with_diff ? 1 : 0) + 1 + 2 + (n >= 2 ? 2 * (2 * n - 1) : 0);
}
/* check_abs_less_than(x [- y], bnd): [x - y]?, t = d + (bnd - 1), check_big_less_than_safe(t, 2*bnd - 1) */
int orc_abs_less_than_witness(const orc_fr *x, const orc_fr *y, size_t count, const uint64_t bnd[4], int lb,
                              orc_fr *out_wit) {
    int W = orc_abs_less_than_witness_count(bnd, lb, y != NULL);
    if (W < 0) return -1;
    uint64_t two[4], bound[4], bm1[4], one[4] = {1, 0, 0, 0};
    u256_add(two, bnd, bnd);
    u256_sub(bound, two, one);
    u256_sub(bm1, bnd, one);
    for (size_t e = 0; e < count; e++) {
        orc_fr *w = out_wit + e * (size_t)W;
        uint64_t d[4], t[4];
        orc_fr_to_canonical(d, &x[e]);
        if (y) {
            uint64_t yy[4];
            orc_fr_to_canonical(yy, &y[e]);
            int_sub_mod(d, d, yy);      /* gate.sub(a, b) */
            w = emit(w, d);
        }
        int_add_mod(t, d, bm1);         /* gate.add(x, bnd - 1) */
        w = emit(w, t);
        w = emit_cbls(w, t, bound, u256_bits(bound), lb);
    }
    return 0;
}
int orc_range_check_witness_count(int range_bits, int lb) {
    if (lb < 1 || lb > 32 || range_bits < 1 || range_bits > 253) return -1;
    int n = ceil_div(range_bits, lb), rem = range_bits % lb;
    return (n >= 2 ? 2 * n - 1 : 0) + (rem > 1 ? 1 : 0);
}
/* RangeChip::range_check(x, range_bits): limbs + running sums, then last_limb * 2^(lb - rem) when rem > 1 */
int orc_range_check_witness(const orc_fr *x, size_t count, int range_bits, int lb, orc_fr *out_wit) {
    int W = orc_range_check_witness_count(range_bits, lb);
    if (W < 0) return -1;
    int n = ceil_div(range_bits, lb), rem = range_bits % lb;
    for (size_t e = 0; e < count; e++) {
        orc_fr *w = out_wit + e * (size_t)W;
        uint64_t v[4];
        orc_fr_to_canonical(v, &x[e]);
        w = emit_range_check(w, v, n, lb);
        if (rem > 1) {
            uint64_t last[4], p2[4];
            orc_fr a, b;
            if (n == 1) {
                memcpy(last, v, sizeof last);
            } else {
                uint64_t t[4];
                u256_shr(t, v, (unsigned)(lb * (n - 1)));
                u256_low_bits(last, t, (unsigned)lb);
            }
            u256_pow2(p2, (unsigned)(lb - rem));
            orc_fr_from_canonical(&a, last);
            orc_fr_from_canonical(&b, p2);
            orc_fr_mul(w, &a, &b);      /* gate.mul(last, 2^(lb - rem)) */
        }
    }
    return 0;
}
/* mat_times_diag_mat: out[i*cols_v + j] = a[i*lda + j] * v[j] */
int orc_mat_times_diag(const orc_fr *a, const orc_fr *v, size_t rows, size_t lda, size_t cols_v, orc_fr *out) {
    if (cols_v > lda) return -1;
    for (size_t i = 0; i < rows; i++)
        for (size_t j = 0; j < cols_v; j++) orc_fr_mul(&out[i * cols_v + j], &a[i * lda + j], &v[j]);
    return 0;
}

/* FixedPointChip041::quantization (SURVEY A.5 / PDF Eq. 11) */
int orc_quantize(const double *x, size_t count, int P, orc_fr *out) {
    if (P < 1 || P > 63) return -1;
    for (size_t i = 0; i < count; i++) {
        /* round half away from zero WITHOUT forming y + 0.5 (which ties-to-even once y >= 2^52) */
        const double y = fabs(x[i]) * ldexp(1.0, P);
        double mag = floor(y);
        if (y - mag >= 0.5) mag += 1.0;
        if (!(mag < ldexp(1.0, 127))) return -2;
        u128 q = (u128)mag;
        uint64_t v[4] = {(uint64_t)q, (uint64_t)(q >> 64), 0, 0};
        if (x[i] < 0 && q != 0) u256_sub(v, MODULUS, v);
        orc_fr_from_canonical(&out[i], v);
    }
    return 0;
}

/* qsqrt model (unpinned): floor(sqrt(a << P)) */
int orc_isqrt_fixed(const orc_fr *a, size_t count, int P, orc_fr *out) {
    for (size_t i = 0; i < count; i++) {
        uint64_t x[4], num[4] = {0, 0, 0, 0}, res[4] = {0, 0, 0, 0}, bit[4], t[4];
        orc_fr_to_canonical(x, &a[i]);
        if (x[2] | x[3]) return -1; /* a must be < 2^128 */
        /* num = x << P (P <= 63) */
        num[0] = x[0] << P;
        num[1] = (x[1] << P) | (P ? x[0] >> (64 - P) : 0);
        num[2] = P ? x[1] >> (64 - P) : 0;
        u256_pow2(bit, 254);
        while (!u256_geq(num, bit) && (bit[0] | bit[1] | bit[2] | bit[3])) u256_shr(bit, bit, 2);
        while (bit[0] | bit[1] | bit[2] | bit[3]) {
            u256_add(t, res, bit);
            u256_shr(res, res, 1);
            if (u256_geq(num, t)) {
                u256_sub(num, num, t);
                u256_add(res, res, bit);
            }
            u256_shr(bit, bit, 2);
        }
        orc_fr_from_canonical(&out[i], res);
    }
    return 0;
}

/* reference src/matrix/mod.rs:299-342 */
int orc_freivalds_witness(const orc_fr *a, const orc_fr *b, const orc_fr *cs, const orc_fr *gamma,
                          size_t n, size_t k, size_t m, orc_fr *powers, orc_fr *prefix_cv,
                          orc_fr *prefix_bv, orc_fr *prefix_abv, orc_fr *diff, orc_fr *is_zero,
                          orc_fr *inv, int threads) {
    if (m < 1) return -1;                                   /* :310 */
    orc_gamma_powers(gamma, m, powers);                     /* :316-326 */
    orc_mat_vec_prefix(cs, powers, n, m, prefix_cv, threads); /* :335 */
    orc_mat_vec_prefix(b, powers, k, m, prefix_bv, threads);  /* :336 */
    /* gather b_times_v = last running sum of each row */
    orc_fr *btv = (orc_fr *)malloc(sizeof(orc_fr) * (k ? k : 1));
    if (!btv) return -2;
    for (size_t i = 0; i < k; i++) btv[i] = prefix_bv[i * m + (m - 1)];
    orc_mat_vec_prefix(a, btv, n, k, prefix_abv, threads);  /* :337 */
    free(btv);
    for (size_t i = 0; i < n; i++) {                        /* :339-341 is_equal */
        orc_fr x = prefix_cv[i * m + (m - 1)];
        orc_fr y = {{0, 0, 0, 0}};
        if (k) y = prefix_abv[i * k + (k - 1)];
        orc_fr_sub(&diff[i], &x, &y);
        if (fr_is_zero(&diff[i])) {
            is_zero[i] = MONT_ONE;
            inv[i] = MONT_ONE;
        } else {
            memset(&is_zero[i], 0, sizeof(orc_fr));
            fr_inv(&inv[i], &diff[i]);
        }
    }
    return 0;
}
