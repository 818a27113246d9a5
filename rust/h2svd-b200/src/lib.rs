//! Rust binding of `include/h2svd_b200.h` and the bulk-assignment shims that replace the value
//! computation of `src/matrix/mod.rs` in neilcouture/halo2-svd041.
//!
//! UNVERIFIED: the build container has no Rust toolchain; this file documents the intended binding
//!.  The cell sequences below are the ones `oracle/pyoracle.py` models and the
//! C++ host mirror (`include/h2svd_zk.hpp`) emits and checks.
#![allow(non_camel_case_types)]
use halo2_base::{
    gates::GateChip,
    utils::BigPrimeField,
    AssignedValue, Context,
    QuantumCell::{Constant, Existing, Witness},
};
use std::os::raw::{c_char, c_int, c_void};

/// bn256::Fr as it sits in memory: 4 x u64 LE limbs, Montgomery form, canonical.
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct h2svd_fr {
    pub l: [u64; 4],
}
#[repr(C)]
pub struct h2svd_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn h2svd_last_error() -> *const c_char;
    pub fn h2svd_create(out: *mut *mut h2svd_ctx, device: c_int, stream: *mut c_void) -> c_int;
    pub fn h2svd_destroy(ctx: *mut h2svd_ctx);
    pub fn h2svd_fr_matmul(ctx: *mut h2svd_ctx, a: *const h2svd_fr, b: *const h2svd_fr, c: *mut h2svd_fr,
                           n: usize, k: usize, m: usize, b_transposed: c_int) -> c_int;
    pub fn h2svd_freivalds_witness(ctx: *mut h2svd_ctx, a: *const h2svd_fr, b: *const h2svd_fr,
                                   c_s: *const h2svd_fr, gamma: *const h2svd_fr, n: usize, k: usize, m: usize,
                                   powers: *mut h2svd_fr, prefix_cv: *mut h2svd_fr, prefix_bv: *mut h2svd_fr,
                                   prefix_abv: *mut h2svd_fr, diff: *mut h2svd_fr, is_zero: *mut h2svd_fr,
                                   inv: *mut h2svd_fr) -> c_int;
    pub fn h2svd_rescale_witness_count(p: c_int, lookup_bits: c_int, shift_bits: c_int, a_num_bits: c_int) -> c_int;
    pub fn h2svd_rescale_witness(ctx: *mut h2svd_ctx, c_s: *const h2svd_fr, count: usize, p: c_int,
                                 lookup_bits: c_int, shift_bits: c_int, a_num_bits: c_int,
                                 out_q: *mut h2svd_fr, out_wit: *mut h2svd_fr) -> c_int;
    pub fn h2svd_zkvec_inner_prefix(ctx: *mut h2svd_ctx, x: *const h2svd_fr, this: *const h2svd_fr,
                                    batch: usize, len: usize, out_prefix: *mut h2svd_fr) -> c_int;
    pub fn h2svd_zkvec_sub(ctx: *mut h2svd_ctx, this: *const h2svd_fr, x: *const h2svd_fr, count: usize,
                           out: *mut h2svd_fr) -> c_int;
    pub fn h2svd_quantize(ctx: *mut h2svd_ctx, x: *const f64, count: usize, p: c_int, out: *mut h2svd_fr) -> c_int;
}

/// Errors keep the reference's behaviour: every `assert!` of src/matrix/mod.rs stays a panic.
fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(h2svd_last_error()) }.to_string_lossy().into_owned();
        panic!("h2svd_b200 error {rc}: {msg}");
    }
}

/// halo2curves Fr <-> wire format: a transmute (same 32 bytes); no conversion.
#[inline]
fn to_wire<F: BigPrimeField>(x: &F) -> h2svd_fr {
    debug_assert_eq!(std::mem::size_of::<F>(), 32);
    unsafe { std::mem::transmute_copy(x) }
}
#[inline]
fn from_wire<F: BigPrimeField>(x: &h2svd_fr) -> F {
    unsafe { std::mem::transmute_copy(x) }
}
fn gather<F: BigPrimeField>(m: &Vec<Vec<AssignedValue<F>>>) -> Vec<h2svd_fr> {
    m.iter().flat_map(|row| row.iter().map(|c| to_wire(c.value()))).collect()
}

pub struct Gpu(*mut h2svd_ctx);
impl Gpu {
    pub fn new(device: i32) -> Self {
        let mut h = std::ptr::null_mut();
        check(unsafe { h2svd_create(&mut h, device, std::ptr::null_mut()) });
        Gpu(h)
    }
}
impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { h2svd_destroy(self.0) }
    }
}

/// Drop-in for `honest_prover_mat_mul` (src/matrix/mod.rs:546-568): same signature plus the handle.
pub fn honest_prover_mat_mul<F: BigPrimeField>(
    gpu: &Gpu, ctx: &mut Context<F>, a: &Vec<Vec<AssignedValue<F>>>, b: &Vec<Vec<AssignedValue<F>>>,
) -> Vec<Vec<AssignedValue<F>>> {
    assert_eq!(a[0].len(), b.len()); // :515
    let (n, k, m) = (a.len(), b.len(), b[0].len());
    let (fa, fb) = (gather(a), gather(b));
    let mut c = vec![h2svd_fr::default(); n * m];
    check(unsafe { h2svd_fr_matmul(gpu.0, fa.as_ptr(), fb.as_ptr(), c.as_mut_ptr(), n, k, m, 0) });
    // :558-565 -- N*M unconstrained witnesses, row-major; assign_witnesses is the bulk form of load_witness
    c.chunks(m).map(|row| ctx.assign_witnesses(row.iter().map(from_wire::<F>))).collect()
}

/// Drop-in for `field_mat_vec_mul` (:574-599) given the GPU-produced running sums of one row:
/// cells `[0, a_0, v_0, s_0, a_1, v_1, s_1, ...]`, gate at every third offset (halo2-base inner_product).
fn assign_inner_product_row<F: BigPrimeField>(
    ctx: &mut Context<F>, a: &[AssignedValue<F>], v: &[AssignedValue<F>], prefix: &[h2svd_fr],
) -> AssignedValue<F> {
    let mut cells = Vec::with_capacity(1 + 3 * a.len());
    cells.push(Constant(F::ZERO));
    for j in 0..a.len() {
        cells.push(Existing(a[j]));
        cells.push(Existing(v[j]));
        cells.push(Witness(from_wire::<F>(&prefix[j])));
    }
    ctx.assign_region_last(cells, (0..a.len()).map(|i| 3 * i as isize))
}

/// Drop-in for `ZkMatrix::verify_mul` (:299-342).  `a`, `b` are the `.matrix` fields.
pub fn verify_mul<F: BigPrimeField>(
    gpu: &Gpu, ctx: &mut Context<F>, gate: &GateChip<F>, a: &Vec<Vec<AssignedValue<F>>>,
    b: &Vec<Vec<AssignedValue<F>>>, c_s: &Vec<Vec<AssignedValue<F>>>, init_rand: &AssignedValue<F>,
) {
    let (n, k, m) = (a.len(), b.len(), b[0].len());
    assert_eq!(a[0].len(), k); // :307
    assert_eq!(c_s.len(), n); // :308
    assert_eq!(c_s[0].len(), m); // :309
    assert!(m >= 1); // :310
    let (fa, fb, fc) = (gather(a), gather(b), gather(c_s));
    let g = to_wire(init_rand.value());
    let z = h2svd_fr::default();
    let (mut pw, mut pcv, mut pbv, mut pabv) = (vec![z; m], vec![z; n * m], vec![z; k * m], vec![z; n * k]);
    let (mut diff, mut isz, mut inv) = (vec![z; n], vec![z; n], vec![z; n]);
    check(unsafe {
        h2svd_freivalds_witness(gpu.0, fa.as_ptr(), fb.as_ptr(), fc.as_ptr(), &g, n, k, m, pw.as_mut_ptr(),
                                pcv.as_mut_ptr(), pbv.as_mut_ptr(), pabv.as_mut_ptr(), diff.as_mut_ptr(),
                                isz.as_mut_ptr(), inv.as_mut_ptr())
    });
    // :318-326  one, then v_i = v_{i-1} * gamma  ([0, v_{i-1}, gamma, v_i], gate at 0)
    let one = ctx.load_witness(from_wire::<F>(&pw[0]));
    gate.assert_is_const(ctx, &one, &F::ONE);
    let mut v = vec![one];
    for i in 1..m {
        let cells = [Constant(F::ZERO), Existing(v[i - 1]), Existing(*init_rand), Witness(from_wire::<F>(&pw[i]))];
        v.push(ctx.assign_region_last(cells, [0]));
    }
    // :335-337
    let cs_v: Vec<_> = (0..n).map(|i| assign_inner_product_row(ctx, &c_s[i], &v, &pcv[i * m..(i + 1) * m])).collect();
    let b_v: Vec<_> = (0..k).map(|i| assign_inner_product_row(ctx, &b[i], &v, &pbv[i * m..(i + 1) * m])).collect();
    let ab_v: Vec<_> = (0..n).map(|i| assign_inner_product_row(ctx, &a[i], &b_v, &pabv[i * k..(i + 1) * k])).collect();
    // :339-341  gate.is_equal = sub (4 cells) + is_zero (8 cells); values come from the GPU
    for i in 0..n {
        let d = ctx.assign_region_last(
            [Witness(from_wire::<F>(&diff[i])), Existing(ab_v[i]), Constant(F::ONE), Existing(cs_v[i])], [0]);
        let (zf, invf): (F, F) = (from_wire(&isz[i]), from_wire(&inv[i]));
        ctx.assign_region(
            [Witness(zf), Existing(d), Witness(invf), Constant(F::ONE), Constant(F::ZERO), Existing(d),
             Witness(zf), Constant(F::ZERO)], [0, 4]);
    }
}

/// Cells of one `range_check(x, n*lb)` given its witnesses `l0, l1, s1, l2, s2, ...` (halo2-base RangeChip):
/// `[l0, l1, 2^lb, s1, l2, 2^2lb, s2, ...]`, gates at 0, 3, ...; every limb is pushed to the lookup table.
fn assign_range_check<F: BigPrimeField>(
    ctx: &mut Context<F>, range: &halo2_base::gates::RangeChip<F>, x: AssignedValue<F>, n: usize, lb: usize, w: &[h2svd_fr],
) -> usize {
    use halo2_base::gates::RangeInstructions;
    if n == 1 {
        range.add_cell_to_lookup(ctx, x, lb); // halo2-base: the single limb is x itself
        return 0;
    }
    let mut cells = vec![Witness(from_wire::<F>(&w[0]))];
    for i in 1..n {
        cells.push(Witness(from_wire::<F>(&w[2 * i - 1])));
        cells.push(Constant(range.gate().pow_of_two()[lb * i]));
        cells.push(Witness(from_wire::<F>(&w[2 * i])));
    }
    let row = ctx.advice.len();
    let acc = ctx.assign_region_last(cells, (0..n - 1).map(|i| 3 * i as isize));
    ctx.constrain_equal(&x, &acc);
    range.add_cell_to_lookup(ctx, ctx.get(row as isize), lb);
    for i in 0..n - 1 {
        range.add_cell_to_lookup(ctx, ctx.get((row + 1 + 3 * i) as isize), lb);
    }
    2 * n - 1
}

/// Cells of one `check_big_less_than_safe(x, bound)` given its witnesses (range_check(x) | chk, xp | range_check(chk)).
fn assign_cbls<F: BigPrimeField>(
    ctx: &mut Context<F>, range: &halo2_base::gates::RangeChip<F>, x: AssignedValue<F>, bound: F, n: usize, lb: usize,
    w: &[h2svd_fr],
) -> usize {
    let mut used = assign_range_check(ctx, range, x, n, lb, w);
    let neg_pow = -range.gate().pow_of_two()[n * lb];
    ctx.assign_region(
        [Witness(from_wire::<F>(&w[used])), Constant(bound), Constant(F::ONE), Witness(from_wire::<F>(&w[used + 1])),
         Constant(neg_pow), Constant(F::ONE), Existing(x)], [0, 3]);
    let chk = ctx.get(-7);
    used += 2;
    used + assign_range_check(ctx, range, chk, n, lb, &w[used..])
}

/// Drop-in for `ZkMatrix::rescale_matrix` (src/matrix/mod.rs:354-375): ONE GPU call for the whole matrix, then the cells of
/// `signed_div_scale` per element in the reference's order.  `shift_bits` / `a_num_bits` are the
/// FixedPointChip041's constants (pass the chip's values; -1 selects 3P / 4P).
pub fn rescale_matrix<F: BigPrimeField>(
    gpu: &Gpu, ctx: &mut Context<F>, range: &halo2_base::gates::RangeChip<F>, c_s: &Vec<Vec<AssignedValue<F>>>,
    precision_bits: u32, lookup_bits: usize, shift_bits: i32, a_num_bits: i32,
) -> Vec<Vec<AssignedValue<F>>> {
    use halo2_base::gates::RangeInstructions;
    let (rows, cols) = (c_s.len(), c_s[0].len());
    let p = precision_bits as usize;
    let s = if shift_bits < 0 { 3 * p } else { shift_bits as usize };
    let a = if a_num_bits < 0 { 4 * p } else { a_num_bits as usize };
    let w_per = unsafe { h2svd_rescale_witness_count(p as c_int, lookup_bits as c_int, s as c_int, a as c_int) };
    assert!(w_per > 0, "rescale parameters out of range");
    let w_per = w_per as usize;
    let (n_d, n_r) = ((a - p + 1 + lookup_bits - 1) / lookup_bits, (p + 1 + lookup_bits - 1) / lookup_bits);
    let fc = gather(c_s);
    let mut q = vec![h2svd_fr::default(); rows * cols];
    let mut wit = vec![h2svd_fr::default(); rows * cols * w_per];
    check(unsafe {
        h2svd_rescale_witness(gpu.0, fc.as_ptr(), rows * cols, p as c_int, lookup_bits as c_int, s as c_int, a as c_int,
                              q.as_mut_ptr(), wit.as_mut_ptr())
    });
    let pow = range.gate().pow_of_two();
    let mut out = Vec::with_capacity(rows);
    for i in 0..rows {
        let mut new_row = Vec::with_capacity(cols);
        for j in 0..cols {
            let w = &wit[(i * cols + j) * w_per..(i * cols + j + 1) * w_per];
            let a_shift = ctx.assign_region_last(
                [Existing(c_s[i][j]), Constant(pow[s]), Constant(F::ONE), Witness(from_wire::<F>(&w[0]))], [0]);
            ctx.assign_region(
                [Witness(from_wire::<F>(&w[1])), Constant(pow[p]), Witness(from_wire::<F>(&w[2])), Existing(a_shift)], [0]);
            let (rem, div) = (ctx.get(-4), ctx.get(-2));
            let mut used = 3;
            used += assign_cbls(ctx, range, div, pow[a - p] + F::ONE, n_d, lookup_bits, &w[used..]);
            used += assign_cbls(ctx, range, rem, pow[p], n_r, lookup_bits, &w[used..]);
            ctx.assign_region(
                [Witness(from_wire::<F>(&w[used])), Constant(pow[s - p]), Constant(F::ONE), Existing(div)], [0]);
            new_row.push(ctx.get(-4));
        }
        out.push(new_row);
    }
    out
}

/// Drop-in for the value side of `ZkVector::inner_product` (:79-106) for ONE pair: running sums from the GPU, cells of
/// `gate.inner_product(u = x, v = self)`; the caller passes the returned cell to `signed_div_scale` (rescale_matrix on a
/// 1x1 matrix, or the chip's own method).
pub fn inner_product_unscaled<F: BigPrimeField>(
    gpu: &Gpu, ctx: &mut Context<F>, this: &Vec<AssignedValue<F>>, x: &Vec<AssignedValue<F>>,
) -> AssignedValue<F> {
    assert!(this.len() == x.len()); // :86
    let fx: Vec<h2svd_fr> = x.iter().map(|c| to_wire(c.value())).collect();
    let fs: Vec<h2svd_fr> = this.iter().map(|c| to_wire(c.value())).collect();
    let mut prefix = vec![h2svd_fr::default(); x.len()];
    check(unsafe { h2svd_zkvec_inner_prefix(gpu.0, fx.as_ptr(), fs.as_ptr(), 1, x.len(), prefix.as_mut_ptr()) });
    assign_inner_product_row(ctx, x, this, &prefix)
}
