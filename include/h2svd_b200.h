/* h2svd_b200 -- C ABI of the B200-native ZkMatrix / ZkVector witness path.
 *
 * This is the drop-in boundary for ONE hot path of neilcouture/halo2-svd041: the value computation
 * of src/matrix/mod.rs (field mat-mul, Freivalds mat-vecs, rescale, ZkVector witnesses).  The
 * reference has no FFI of its own (pure Rust); these entry points are what a Rust `extern "C"`
 * block in the reference's src/matrix would bind (see INTEGRATION.md for the stub).  Every entry
 * point cites the reference function (file:line under /root/reference) whose value computation it
 * replaces.
 *
 * Conventions
 *  - Field elements are halo2curves bn256::Fr exactly as they sit in Rust memory: 32 bytes,
 *    4 x u64 little-endian limbs of x*2^256 mod r (Montgomery form), canonical (< r).  Arrays are
 *    row-major and contiguous.  No conversion happens on either side of the boundary.
 *  - Every function returns 0 on success or a negative H2SVD_E* code; nothing unwinds, nothing
 *    aborts.  h2svd_last_error() returns a static/thread-local description of the last failure.
 *  - A handle owns one CUDA device, one stream (plus two internal side streams) and grow-only device workspaces.  It
 *    is NOT thread-safe (it mirrors the reference's `&mut Context`): one handle per thread / per GPU.  There is no
 *    process-global mutable state: two handles never influence each other.  h2svd_multi (below) bundles one handle
 *    per GPU for single-process multi-GPU callers.
 *  - Functions without suffix take HOST pointers (pageable or pinned) and do H2D + kernels + D2H
 *    before returning.  Functions ending in _dev take DEVICE pointers on the handle's device, are
 *    asynchronous on the handle's stream (h2svd_sync to wait) and never touch host memory.
 *  - There is no CPU fallback: without a usable sm_100-class GPU h2svd_create fails.
 */
#ifndef H2SVD_B200_H
#define H2SVD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h2svd_fr { uint64_t l[4]; } h2svd_fr;      /* bn256::Fr, Montgomery, canonical */
typedef struct h2svd_ctx h2svd_ctx;                       /* opaque handle */

enum {
    H2SVD_OK = 0,
    H2SVD_EINVAL = -1,    /* bad argument (null pointer, shape mismatch, parameter out of range) */
    H2SVD_ECUDA = -2,     /* CUDA runtime error (message in h2svd_last_error) */
    H2SVD_ENOMEM = -3,    /* device or host allocation failed */
    H2SVD_ENODEV = -4,    /* no usable GPU: there is no CPU fallback */
    H2SVD_ERANGE = -5     /* a non-canonical field element (>= r) was found in an input */
};

const char *h2svd_last_error(void);
const char *h2svd_version(void);

/* ---- handle ------------------------------------------------------------------------------------ */
/* device < 0 selects the current CUDA device.  `stream` is a cudaStream_t to adopt (e.g. a torch
 * stream) or NULL to let the handle create its own non-blocking stream. */
int h2svd_create(h2svd_ctx **out, int device, void *stream);
void h2svd_destroy(h2svd_ctx *ctx);
int h2svd_sync(h2svd_ctx *ctx);
void *h2svd_stream(h2svd_ctx *ctx);
int h2svd_device(h2svd_ctx *ctx);
int h2svd_sm_count(h2svd_ctx *ctx);
/* Number of kernel launches issued through this handle since creation (bench.py's gpu_launches). */
uint64_t h2svd_launch_count(h2svd_ctx *ctx);

/* ---- K1: field mat-mul ---------------------------------------------------------------------------
 * Replaces field_mat_mul (src/matrix/mod.rs:510-537), the O(N^3) loop inside
 * honest_prover_mat_mul (:546-568).  c[n x m] = a[n x k] * b[k x m] over Fr.  With
 * b_transposed != 0, `b` holds the m x k matrix B^T (what ZkMatrix::transpose_matrix, :408, would
 * have been called on), so callers such as check_svd_phase0 (src/svd/mod.rs:96,109,112) need not
 * materialise the transpose.  Asserts of the reference (:515) become H2SVD_EINVAL.
 * Several engines compute the same bytes: from 64^3 on (and k >= 32) the product runs on the tensor cores as exact 8-bit
 * integer MMAs (csrc/matmul_tc.cu) -- over 9 x 9 signed byte digits when every operand is a small signed integer in
 * standard form (|x| < 2^70: what ZkMatrix::new's quantization produces; detected on the device), else over the 32 x 32
 * byte planes of the full-width Montgomery representation -- smaller products on the integer pipe (csrc/matmul.cu).  The
 * choice is internal; results do not depend on it. */
int h2svd_fr_matmul(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *c, size_t n,
                    size_t k, size_t m, int b_transposed);
int h2svd_fr_matmul_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *c,
                        size_t n, size_t k, size_t m, int b_transposed);

/* ---- K2/K3: Freivalds witness ----------------------------------------------------------------------
 * Replaces the value computation of ZkMatrix::verify_mul (src/matrix/mod.rs:299-342):
 *   powers[m]       v_i = gamma^i               (:316-326; v_0 = 1 is the `one` witness of :318)
 *   prefix_cv[n*m]  running sums of c_s . v     (:335 -> field_mat_vec_mul :574-599)
 *   prefix_bv[k*m]  running sums of b . v       (:336)
 *   prefix_abv[n*k] running sums of a . (b v)   (:337)
 *   diff[n], is_zero[n], inv[n]                 the Witness cells of gate.is_equal (:339-341):
 *                                               diff = (c_s v)_i - (a b v)_i, is_zero in {0,1},
 *                                               inv = 1 if diff == 0 else diff^-1
 * i.e. every `Witness`-kind advice value of verify_mul; Existing/Constant cells are re-emitted by
 * the caller from its own inputs.  Shape asserts (:307-310) -> H2SVD_EINVAL. */
int h2svd_freivalds_witness(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b,
                            const h2svd_fr *c_s, const h2svd_fr *gamma, size_t n, size_t k,
                            size_t m, h2svd_fr *powers, h2svd_fr *prefix_cv, h2svd_fr *prefix_bv,
                            h2svd_fr *prefix_abv, h2svd_fr *diff, h2svd_fr *is_zero,
                            h2svd_fr *inv);
int h2svd_freivalds_witness_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b,
                                const h2svd_fr *c_s, const h2svd_fr *gamma, size_t n, size_t k,
                                size_t m, h2svd_fr *powers, h2svd_fr *prefix_cv,
                                h2svd_fr *prefix_bv, h2svd_fr *prefix_abv, h2svd_fr *diff,
                                h2svd_fr *is_zero, h2svd_fr *inv);

/* Building blocks of the above, exposed for row-sharded (multi-GPU) callers:
 * gamma powers (:316-326) and one prefix-sum mat-vec (field_mat_vec_mul :574-599). */
int h2svd_gamma_powers_dev(h2svd_ctx *ctx, const h2svd_fr *gamma, size_t d, h2svd_fr *out);
int h2svd_mat_vec_prefix_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows,
                             size_t len, h2svd_fr *out_prefix);
/* Same, additionally writing out_totals[i] = out_prefix[i*len + len-1] (the value field_mat_vec_mul
 * returns for row i, :597); out_totals may be NULL. */
int h2svd_mat_vec_prefix_totals_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows,
                                    size_t len, h2svd_fr *out_prefix, h2svd_fr *out_totals);
/* Two matrices against the same vector in ONE launch (c_s . v and b . v of verify_mul, :335-336);
 * rows1 may be 0. */
int h2svd_mat_vec_prefix_pair_dev(h2svd_ctx *ctx, const h2svd_fr *a0, size_t rows0,
                                  h2svd_fr *out_prefix0, h2svd_fr *out_totals0, const h2svd_fr *a1,
                                  size_t rows1, h2svd_fr *out_prefix1, h2svd_fr *out_totals1,
                                  const h2svd_fr *v, size_t len);
/* Row totals only: out_totals[i] = sum_t a[i*len + t] * v[t] -- the value field_mat_vec_mul returns per row (:597) without
 * the running-sum witnesses (lazy accumulation, one reduction per row).  Row-sharded callers use it for (b v). */
int h2svd_mat_vec_totals_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t len,
                             h2svd_fr *out_totals);
/* out[i] = src[i*stride + offset]  (gathers the last running sum of every row) */
int h2svd_gather_dev(h2svd_ctx *ctx, const h2svd_fr *src, size_t count, size_t stride,
                     size_t offset, h2svd_fr *out);
/* is_equal witness cells (:339-341) for two vectors of row totals */
int h2svd_is_equal_witness_dev(h2svd_ctx *ctx, const h2svd_fr *x, const h2svd_fr *y, size_t count,
                               h2svd_fr *diff, h2svd_fr *is_zero, h2svd_fr *inv);

/* Host-pointer form of h2svd_mat_vec_prefix_dev: the value side of field_mat_vec_mul (:574-599) for callers outside
 * verify_mul (e.g. ZkVector::mul, :169-182).  out_prefix[i*len + j] = sum_{t<=j} a[i][t] * v[t]. */
int h2svd_mat_vec_prefix(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t len,
                         h2svd_fr *out_prefix);

/* ---- K4: rescale witness -----------------------------------------------------------------------------
 * Replaces the value computation of ZkMatrix::rescale_matrix (src/matrix/mod.rs:354-375), i.e. one
 * FixedPointChip041::signed_div_scale per element (:369; also ZkVector::inner_product :104).
 * For each of `count` elements: out_q = floor(a_signed / 2^P) as a field element (the returned
 * matrix entry), and out_wit[e*W .. (e+1)*W) = the W `Witness`-kind advice values of that call in
 * assignment order (a_shift, rem, div, 4 limb-decomposition blocks, q).
 * shift_bits / a_num_bits are the chip's constants (third-party, unpinned): pass -1 for the
 * defaults 3P / 4P.  h2svd_rescale_witness_count returns W (= 4 + f(n_d) + f(n_r) with f(n) = 4n, or 2 when
 * n == 1: a one-limb range check emits no limb cells) or <0. */
int h2svd_rescale_witness_count(int precision_bits, int lookup_bits, int shift_bits,
                                int a_num_bits);
int h2svd_rescale_witness(h2svd_ctx *ctx, const h2svd_fr *c_s, size_t count, int precision_bits,
                          int lookup_bits, int shift_bits, int a_num_bits, h2svd_fr *out_q,
                          h2svd_fr *out_wit);
int h2svd_rescale_witness_dev(h2svd_ctx *ctx, const h2svd_fr *c_s, size_t count,
                              int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                              h2svd_fr *out_q, h2svd_fr *out_wit);

/* honest_prover_mat_mul followed by rescale_matrix of the product (the README.md:34-47 sequence; src/matrix/mod.rs:546
 * then :354): c_s = A.B, out_q / out_wit exactly as h2svd_rescale_witness_dev(c_s, n*m, ...) would fill them.  Runs the
 * two kernels back to back (an experimental fused launch exists behind a tuning hook; same bytes).  Device pointers. */
int h2svd_fr_matmul_rescale_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b, size_t n, size_t k,
                                size_t m, int precision_bits, int lookup_bits, int shift_bits,
                                int a_num_bits, h2svd_fr *c_s, h2svd_fr *out_q, h2svd_fr *out_wit);

/* ---- range-check witnesses of the SVD verifier's helpers ---------------------------------------------
 * h2svd_abs_less_than_witness: check_abs_less_than (src/matrix/mod.rs:425-437) for `count` elements, optionally of a
 *   difference (check_mat_diff :441-459, check_mat_id :461-483; pass y = NULL for check_mat_entries_bounded :490-501):
 *     d = x - y                                  gate.sub, 1 Witness            (only when y != NULL)
 *     t = d + (bnd - 1)                          gate.add, 1 Witness
 *     check_big_less_than_safe(t, 2*bnd - 1)     range_check(t) | chk, xp | range_check(chk)   
 *   `bnd` is the canonical integer bound, 4 x u64 little-endian.  out_wit[count * W], W from the _count function.
 * h2svd_range_check_witness: RangeChip::range_check(x, range_bits) (ZkVector::entries_less_than :185-197,
 *   entries_in_desc_order :199-216): n = ceil(range_bits/lookup_bits) limbs and n-1 running sums (l0, l1, s1, l2, s2, ...;
 *   none when n == 1), then last_limb * 2^(lookup_bits - rem) when range_bits % lookup_bits = rem > 1.  W may be 0.
 * h2svd_mat_times_diag: mat_times_diag_mat (:610-627): out[i*cols_v + j] = a[i*lda + j] * v[j] (the gate.mul Witness). */
int h2svd_abs_less_than_witness_count(const uint64_t bnd[4], int lookup_bits, int with_diff);
int h2svd_abs_less_than_witness(h2svd_ctx *ctx, const h2svd_fr *x, const h2svd_fr *y, size_t count,
                                const uint64_t bnd[4], int lookup_bits, h2svd_fr *out_wit);
int h2svd_abs_less_than_witness_dev(h2svd_ctx *ctx, const h2svd_fr *x, const h2svd_fr *y, size_t count,
                                    const uint64_t bnd[4], int lookup_bits, h2svd_fr *out_wit);
int h2svd_range_check_witness_count(int range_bits, int lookup_bits);
int h2svd_range_check_witness(h2svd_ctx *ctx, const h2svd_fr *x, size_t count, int range_bits, int lookup_bits,
                              h2svd_fr *out_wit);
int h2svd_range_check_witness_dev(h2svd_ctx *ctx, const h2svd_fr *x, size_t count, int range_bits,
                                  int lookup_bits, h2svd_fr *out_wit);
int h2svd_mat_times_diag(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t lda,
                         size_t cols_v, h2svd_fr *out);
int h2svd_mat_times_diag_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *v, size_t rows, size_t lda,
                             size_t cols_v, h2svd_fr *out);

/* ---- the whole README.md:34-47 sequence in one pipelined call -----------------------------------------------------
 * honest_prover_mat_mul (src/matrix/mod.rs:546) -> rescale_matrix (:354) -> verify_mul (:299) for the caller's `rows`
 * rows of A (all of A on one GPU; a row slab when the job is sharded over several handles/GPUs) against all of B.
 * Host pointers in, host pointers out; internally the rows are processed in slabs so that the device-to-host copy of
 * slab s (the rescale witnesses dominate: rows*m*W*32 bytes) overlaps the kernels of slab s+1 -- the call is bound by
 * the PCIe transfer of its outputs, not by the sum of transfer and compute.  Use h2svd_host_alloc'ed (pinned) buffers
 * for full bandwidth; pageable memory works but is staged by the driver.
 *   c_s[rows*m]              the unscaled product (the Witness cells of honest_prover_mat_mul)
 *   q[rows*m], wit[rows*m*W] as h2svd_rescale_witness
 *   powers[m], prefix_cv[rows*m], prefix_abv[rows*k], diff/is_zero/inv[rows]   as h2svd_freivalds_witness
 *   prefix_bv[(bv_row1-bv_row0)*m]  running sums of rows [bv_row0, bv_row1) of b . v (pass 0, k for all of them; a
 *                                   sharded caller asks each handle for a different range -- no exchange needed,
 *                                   every handle computes the k row totals of b . v itself) */
int h2svd_zkmatrix_mul_witness(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b, const h2svd_fr *gamma,
                               size_t rows, size_t k, size_t m, int precision_bits, int lookup_bits,
                               int shift_bits, int a_num_bits, size_t bv_row0, size_t bv_row1, h2svd_fr *c_s,
                               h2svd_fr *q, h2svd_fr *wit, h2svd_fr *powers, h2svd_fr *prefix_cv,
                               h2svd_fr *prefix_bv, h2svd_fr *prefix_abv, h2svd_fr *diff, h2svd_fr *is_zero,
                               h2svd_fr *inv);

/* Device-pointer form of the same sequence (what h2svd_zkmatrix_mul_witness runs per slab, and what a caller that keeps
 * its matrices on the GPU uses): asynchronous on the handle's stream.  Internally the C-independent half of verify_mul
 * (gamma powers :316-326, b . v :336, a . (b v) :337 -- integer-pipe work) is forked onto a side stream under the mat-mul
 * (tensor pipe), and c_s . v (:335) + is_equal (:339-341) run next to the rescale kernel (HBM writes); the side stream is
 * joined before the call returns to the caller's stream order.  prefix_bv receives rows [bv_row0, bv_row1) of b . v; every
 * handle derives all k row totals itself, so row-sharded callers need no exchange step. */
int h2svd_zkmatrix_mul_witness_dev(h2svd_ctx *ctx, const h2svd_fr *a, const h2svd_fr *b, const h2svd_fr *gamma,
                                   size_t rows, size_t k, size_t m, int precision_bits, int lookup_bits,
                                   int shift_bits, int a_num_bits, size_t bv_row0, size_t bv_row1, h2svd_fr *c_s,
                                   h2svd_fr *q, h2svd_fr *wit, h2svd_fr *powers, h2svd_fr *prefix_cv,
                                   h2svd_fr *prefix_bv, h2svd_fr *prefix_abv, h2svd_fr *diff, h2svd_fr *is_zero,
                                   h2svd_fr *inv);

/* ---- CUDA-graph capture -------------------------------------------------------------------------------------------
 * Any sequence of *_dev calls on one handle between h2svd_graph_begin and h2svd_graph_end is recorded instead of run
 * (cudaStreamBeginCapture on the handle's stream) and can then be replayed with ONE launch: a row slab on one of 8 GPUs
 * is ~20 kernels of 2-70 us each, where launch gaps are a visible share of the step.  Rules: run the same calls once
 * un-captured first (workspaces only grow outside a capture); the recorded pointers and shapes are baked in -- if a larger
 * call makes a workspace of the handle grow afterwards, h2svd_graph_launch refuses the stale graph (H2SVD_EINVAL: record it
 * again); host-pointer entry points cannot be captured.  h2svd_launch_count advances by the recorded kernel count per replay. */
typedef struct h2svd_graph h2svd_graph;
int h2svd_graph_begin(h2svd_ctx *ctx);
int h2svd_graph_end(h2svd_ctx *ctx, h2svd_graph **out);
int h2svd_graph_launch(h2svd_ctx *ctx, h2svd_graph *graph);
void h2svd_graph_destroy(h2svd_graph *graph);

/* ---- one process, several GPUs (north_star: rows of A and C shard across the GPUs of one box, B replicated) ---------
 * h2svd_multi owns one handle per entry of devices[] (an index may repeat: two handles on one GPU exercise the same
 * partitioning).  h2svd_multi_zkmatrix_mul_witness is h2svd_zkmatrix_mul_witness for ALL n rows: handle g gets the
 * contiguous row range g of A / C (first n % parts ranges one row longer) and the matching range of the rows of b . v,
 * the per-handle calls run concurrently (one host thread per GPU, each bound by its own PCIe link), results land in the
 * caller's arrays at the right offsets.  No data-path exchange between GPUs: every handle derives (b v) itself, and
 * field addition is exact, so the output is byte-identical to the single-GPU call.  This is what a caller of
 * ZkMatrix::verify_mul (src/matrix/mod.rs:299) binds to use the whole box. */
typedef struct h2svd_multi h2svd_multi;
int h2svd_multi_create(h2svd_multi **out, const int *devices, int n_dev);
void h2svd_multi_destroy(h2svd_multi *mh);
int h2svd_multi_count(h2svd_multi *mh);
h2svd_ctx *h2svd_multi_ctx(h2svd_multi *mh, int i);
int h2svd_multi_zkmatrix_mul_witness(h2svd_multi *mh, const h2svd_fr *a, const h2svd_fr *b, const h2svd_fr *gamma,
                                     size_t n, size_t k, size_t m, int precision_bits, int lookup_bits, int shift_bits,
                                     int a_num_bits, h2svd_fr *c_s, h2svd_fr *q, h2svd_fr *wit, h2svd_fr *powers,
                                     h2svd_fr *prefix_cv, h2svd_fr *prefix_bv, h2svd_fr *prefix_abv, h2svd_fr *diff,
                                     h2svd_fr *is_zero, h2svd_fr *inv);

/* Pinned (page-locked) host memory for the host-pointer entry points. */
int h2svd_host_alloc(size_t bytes, void **out);
void h2svd_host_free(void *p);

/* ---- K5/K6: ZkVector witnesses -------------------------------------------------------------------------
 * h2svd_zkvec_inner_prefix: running sums of gate.inner_product(u = x, v = self) for `batch`
 *   independent vector pairs (ZkVector::inner_product, src/matrix/mod.rs:79-100; _norm_square :111
 *   is the x == self case).  out_prefix[b*len + j] = sum_{t<=j} x[b][t]*self[b][t]; the caller feeds
 *   out_prefix[b*len + len-1] to h2svd_rescale_witness for the signed_div_scale of :104.
 * h2svd_zkvec_sub: diff[i] = self[i] - x[i], the Witness cell of each fpchip.qsub in
 *   ZkVector::_dist_square (:143-146).
 * h2svd_isqrt_fixed: value model of fpchip.qsqrt (:130, :163): floor(sqrt(a * 2^P)) for a < 2^128
 *   (third-party semantics, parity unpinned -- SURVEY.md A.6). */
int h2svd_zkvec_inner_prefix(h2svd_ctx *ctx, const h2svd_fr *x, const h2svd_fr *self, size_t batch,
                             size_t len, h2svd_fr *out_prefix);
int h2svd_zkvec_inner_prefix_dev(h2svd_ctx *ctx, const h2svd_fr *x, const h2svd_fr *self,
                                 size_t batch, size_t len, h2svd_fr *out_prefix);
int h2svd_zkvec_sub(h2svd_ctx *ctx, const h2svd_fr *self, const h2svd_fr *x, size_t count,
                    h2svd_fr *out);
int h2svd_zkvec_sub_dev(h2svd_ctx *ctx, const h2svd_fr *self, const h2svd_fr *x, size_t count,
                        h2svd_fr *out);
int h2svd_isqrt_fixed(h2svd_ctx *ctx, const h2svd_fr *a, size_t count, int precision_bits,
                      h2svd_fr *out);
int h2svd_isqrt_fixed_dev(h2svd_ctx *ctx, const h2svd_fr *a, size_t count, int precision_bits,
                          h2svd_fr *out);

/* ---- quantization ------------------------------------------------------------------
 * FixedPointChip041::quantization as used by ZkMatrix::new / ZkVector::new
 * (src/matrix/mod.rs:29-40, :230-252): f64 -> Fr, sign-magnitude round-half-up of |x|*2^P,
 * negatives as r - q. */
int h2svd_quantize(h2svd_ctx *ctx, const double *x, size_t count, int precision_bits,
                   h2svd_fr *out);
int h2svd_quantize_dev(h2svd_ctx *ctx, const double *x, size_t count, int precision_bits,
                       h2svd_fr *out);

/* ---- bulk hand-off of the witnesses into halo2-base ---------------------------------------------------
 * The reference assigns ~100 advice cells per rescaled element one ctx.assign_region at a time (src/matrix/mod.rs:364-373
 * -> signed_div_scale -> RangeChip), and the proving backend then walks the Context's advice vector
 * (src/utils/executor.rs:100,116).  Every unit (element / row) of one operation has the same cell structure, so:
 *   h2svd_*_cells_layout     the static structure of ONE unit: per cell its kind and where its value comes from, the
 *                            offsets with the gate selector on, the cells pushed to the lookup table (in push order), the
 *                            extra constrain_equal pairs, the constants.  Offsets >= 0 are cells of the unit,
 *                            -1 - i is the unit's i-th input cell (an AssignedValue the caller already holds).
 *   h2svd_expand_cells       the VALUES of all units in assignment order, out_values[u * cells + c], produced from the
 *                            Witness-only array of the matching h2svd_*_witness call (and the inputs) in one pass of
 *                            32-byte copies on `threads` host threads (0 = all).
 * A binding appends out_values to Context::advice with one extend and replays the layout per unit for selectors, lookups
 * and copy constraints.  Host functions: no GPU, no handle. */
enum { H2SVD_CELL_WITNESS = 0, H2SVD_CELL_CONSTANT = 1, H2SVD_CELL_EXISTING = 2 };
typedef struct h2svd_cells_layout {
    uint32_t cells;        /* advice cells per unit */
    uint32_t witnesses;    /* Witness values per unit == W of the matching h2svd_*_witness call */
    uint32_t inputs;       /* input cells per unit */
    uint32_t n_gates, n_lookups, n_copies, n_constants;
    const uint8_t *kind;          /* [cells] H2SVD_CELL_* */
    const int32_t *source;        /* [cells] Witness: index into the unit's witness stripe; Constant: index into constants;
                                     Existing: cell of the unit (>= 0, always earlier) or input (-1 - i) */
    const uint32_t *gates;        /* [n_gates] offsets q with the vertical gate a + b*c - d == 0 on cells q .. q+3 */
    const int32_t *lookups;       /* [n_lookups] cells constrained to [0, 2^lookup_bits), in push order */
    const int32_t *copies;        /* [2 * n_copies] constrain_equal(a, b) pairs beyond those implied by Existing cells */
    const h2svd_fr *constants;    /* [n_constants] */
} h2svd_cells_layout;
/* FixedPointChip041::signed_div_scale per element (rescale_matrix :354-375, inner_product :104); input 0 = the element */
int h2svd_rescale_cells_layout(int precision_bits, int lookup_bits, int shift_bits, int a_num_bits,
                               h2svd_cells_layout **out);
/* check_abs_less_than(x [- y], bnd) (:425-459); input 0 = x, input 1 = y when with_diff */
int h2svd_abs_less_than_cells_layout(const uint64_t bnd[4], int lookup_bits, int with_diff, h2svd_cells_layout **out);
/* RangeChip::range_check(x, range_bits) (:185-216); input 0 = x */
int h2svd_range_check_cells_layout(int range_bits, int lookup_bits, h2svd_cells_layout **out);
/* gate.is_equal(a, b) of verify_mul (:339-341); inputs a, b; witness stripe diff, is_zero, inv */
int h2svd_is_equal_cells_layout(h2svd_cells_layout **out);
void h2svd_cells_layout_destroy(h2svd_cells_layout *layout);
int h2svd_expand_cells(const h2svd_cells_layout *layout, const h2svd_fr *inputs, const h2svd_fr *wit, size_t units,
                       h2svd_fr *out_values, int threads);
/* field_mat_vec_mul rows (:574-599): [0, a_0, v_0, s_0, a_1, v_1, s_1, ...], 1 + 3*len cells per row, gates at 3*j;
 * v_row_stride = 0 for one shared vector, len for a vector per row (ZkVector::inner_product, u = x, v = self). */
int h2svd_expand_inner_product_cells(const h2svd_fr *a, const h2svd_fr *v, size_t v_row_stride, const h2svd_fr *prefix,
                                     size_t rows, size_t len, h2svd_fr *out_values, int threads);
/* the challenge powers of verify_mul (:316-326): [one] + (m-1) x [0, v_{i-1}, gamma, v_i]; 1 + 4*(m-1) cells */
int h2svd_expand_gamma_power_cells(const h2svd_fr *gamma, const h2svd_fr *powers, size_t m, h2svd_fr *out_values);
/* gate.is_equal per row from the separate arrays of h2svd_freivalds_witness: 12 cells per row */
int h2svd_expand_is_equal_cells(const h2svd_fr *x, const h2svd_fr *y, const h2svd_fr *diff, const h2svd_fr *is_zero,
                                const h2svd_fr *inv, size_t count, h2svd_fr *out_values);

/* ---- input validation ----------------------------------------------------------------------------------------
 * Returns H2SVD_OK if all `count` device-resident elements are canonical (< r), else H2SVD_ERANGE. */
int h2svd_check_canonical_dev(h2svd_ctx *ctx, const h2svd_fr *x, size_t count);

/* ---- host-side scalar field helpers (no GPU, no handle) --------------------------------------------------------------
 * What a shim needs to build `Constant` cells (2^S, 2^(lb*i), -2^bits, ...) and to check gates a + b*c - d == 0
 * on an advice stream; scalar utilities, not a compute path.  from_canonical returns H2SVD_ERANGE for x >= r. */
int h2svd_host_fr_from_canonical(const uint64_t x[4], h2svd_fr *out);
void h2svd_host_fr_to_canonical(const h2svd_fr *a, uint64_t out[4]);
void h2svd_host_fr_add(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *out);
void h2svd_host_fr_sub(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *out);
void h2svd_host_fr_mul(const h2svd_fr *a, const h2svd_fr *b, h2svd_fr *out);

/* ---- measurement aids -------------------------------------------------------------------------------------------
 * Integer-pipe micro-benchmark used as the mat-mul roofline denominator.  kind: 0 = mad.lo.u32
 * (IMAD), 1 = mad.wide.u32 (IMAD.WIDE.U32), 2 = IMAD.WIDE.U32.X carry chains exactly as the mat-mul
 * inner loop issues them, 3 = full 8x8 lazy multiply-accumulate (64 IMAD.WIDE + 16 IADD3.X).
 * Writes achieved multiply instructions per second (thread-level ops) to *ops_per_s. */
int h2svd_microbench_imad(h2svd_ctx *ctx, int kind, int iters, double *ops_per_s);
/* HBM micro-benchmark over `bytes` of the handle's workspace, best of five launches after a warm-up: kind 0 = copy (read +
 * write, what MEASURED_PEAKS.json hbm_gbs measures), 1 = write-only streaming 16-byte stores, 2 = write-only 256-byte bulk
 * stores from shared memory (the witness stream's path), 3 = read-only.  The witness kernels are 97 % writes: their
 * roofline is the WRITE figure, reported beside the copy figure.  Writes GB/s (bytes moved / time). */
int h2svd_microbench_hbm(h2svd_ctx *ctx, int kind, size_t bytes, double *gb_per_s);
/* Tensor-pipe micro-benchmark, the measured denominator of the tensor-core mat-mul engines: back-to-back
 * tcgen05.mma.kind::i8 of the real kernels' shape with operands resident in shared memory, one CTA per SM.
 * kind 0 = unsigned, M128 N256 K32 (full-width engine); 1 = signed, M128 N256 K32 (small-operand engine, 28-column tiles);
 * 2, 3 = signed, N144 / N80 (its 16- and 8-column tiles: the time of one MMA does not shrink with N).
 * min_seconds <= 0: best of three ~4 ms launches (burst); > 0: launches back to back for at least that long (sustained
 * under the power cap).  Writes 8-bit ops per second (multiply-add = 2 ops). */
int h2svd_microbench_tensor_i8(h2svd_ctx *ctx, int kind, double min_seconds, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* H2SVD_B200_H */
