//! Rust binding of `include/h2svd_b200.h` and the drop-in `ZkMatrix` / `ZkVector` of neilcouture/halo2-svd041
//! (`src/matrix/mod.rs`) whose value computation runs on the B200 library.
//!
//! UNVERIFIED: the build container has no Rust toolchain (no cargo, no crates); this file is the binding a maintainer
//! adds.  What IS built and tested in this repo: the C ABI every `extern "C"` item below names, the
//! cell layouts (`h2svd_*_cells_layout`, diffed against the halo2-base model in tests/test_expand_cells.py), and a C++
//! mirror of these shims (`include/h2svd_zk.hpp`) whose advice streams are compared cell by cell with the oracle.
//!
//! Every `impl` item keeps the reference's name, argument list and panics; file:line cites the reference.
#![allow(non_camel_case_types, clippy::needless_return)]
use halo2_base::{
    gates::{GateChip, GateInstructions, RangeChip, RangeInstructions},
    utils::BigPrimeField,
    AssignedValue, Context,
    QuantumCell::{Constant, Existing, Witness},
};
use std::cell::RefCell;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::atomic::{AtomicI32, Ordering};
use zk_fixed_point_chip::gadget::fixed_point041::{FixedPointChip041, FixedPointInstructions041};

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
#[repr(C)]
pub struct h2svd_multi {
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
    pub fn h2svd_mat_vec_prefix(ctx: *mut h2svd_ctx, a: *const h2svd_fr, v: *const h2svd_fr, rows: usize, len: usize,
                                out_prefix: *mut h2svd_fr) -> c_int;
    pub fn h2svd_rescale_witness_count(p: c_int, lookup_bits: c_int, shift_bits: c_int, a_num_bits: c_int) -> c_int;
    pub fn h2svd_rescale_witness(ctx: *mut h2svd_ctx, c_s: *const h2svd_fr, count: usize, p: c_int,
                                 lookup_bits: c_int, shift_bits: c_int, a_num_bits: c_int,
                                 out_q: *mut h2svd_fr, out_wit: *mut h2svd_fr) -> c_int;
    pub fn h2svd_zkvec_inner_prefix(ctx: *mut h2svd_ctx, x: *const h2svd_fr, this: *const h2svd_fr,
                                    batch: usize, len: usize, out_prefix: *mut h2svd_fr) -> c_int;
    pub fn h2svd_zkvec_sub(ctx: *mut h2svd_ctx, this: *const h2svd_fr, x: *const h2svd_fr, count: usize,
                           out: *mut h2svd_fr) -> c_int;
    pub fn h2svd_quantize(ctx: *mut h2svd_ctx, x: *const f64, count: usize, p: c_int, out: *mut h2svd_fr) -> c_int;
    /// The README.md:34-47 sequence (mat-mul -> rescale -> verify_mul) for all n rows over every GPU of the box.
    pub fn h2svd_multi_create(out: *mut *mut h2svd_multi, devices: *const c_int, n_dev: c_int) -> c_int;
    pub fn h2svd_multi_destroy(mh: *mut h2svd_multi);
    pub fn h2svd_multi_zkmatrix_mul_witness(mh: *mut h2svd_multi, a: *const h2svd_fr, b: *const h2svd_fr,
                                            gamma: *const h2svd_fr, n: usize, k: usize, m: usize, p: c_int, lookup_bits: c_int,
                                            shift_bits: c_int, a_num_bits: c_int, c_s: *mut h2svd_fr, q: *mut h2svd_fr,
                                            wit: *mut h2svd_fr, powers: *mut h2svd_fr, prefix_cv: *mut h2svd_fr,
                                            prefix_bv: *mut h2svd_fr, prefix_abv: *mut h2svd_fr, diff: *mut h2svd_fr,
                                            is_zero: *mut h2svd_fr, inv: *mut h2svd_fr) -> c_int;
}

/// Errors keep the reference's behaviour: every `assert!` of src/matrix/mod.rs stays a panic.
fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(h2svd_last_error()) }.to_string_lossy().into_owned();
        panic!("h2svd_b200 error {rc}: {msg}");
    }
}

/// `FixedPointChip041::signed_div_scale`'s constants are not pinned by the reference (git dependency without a rev,
/// Cargo.toml:64): shift S and a_num_bits A default to 3P / 4P (-1) and can be set once to the chip's values.
pub static SHIFT_BITS: AtomicI32 = AtomicI32::new(-1);
pub static A_NUM_BITS: AtomicI32 = AtomicI32::new(-1);

/// One handle per thread, created on first use on device `H2SVD_DEVICE` (default 0) -- the reference's methods take no
/// handle, and a `Context` is `&mut` (single-threaded) anyway.  Mirrors `Gpu::current()` of include/h2svd_zk.hpp.
struct Gpu(*mut h2svd_ctx);
impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { h2svd_destroy(self.0) }
    }
}
thread_local! {
    static GPU: RefCell<Option<Gpu>> = RefCell::new(None);
}
fn gpu() -> *mut h2svd_ctx {
    GPU.with(|g| {
        let mut g = g.borrow_mut();
        if g.is_none() {
            let dev = std::env::var("H2SVD_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let mut h = std::ptr::null_mut();
            check(unsafe { h2svd_create(&mut h, dev, std::ptr::null_mut()) });
            *g = Some(Gpu(h));
        }
        g.as_ref().unwrap().0
    })
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
fn gather1<F: BigPrimeField>(v: &[AssignedValue<F>]) -> Vec<h2svd_fr> {
    v.iter().map(|c| to_wire(c.value())).collect()
}

// ---------------------------------------------------------------------------------------------------------------------
// cell layouts of halo2-base 0.4.1 with GPU-supplied values (the same sequences include/h2svd_zk.hpp emits and
// oracle/pyoracle.py models; h2svd_*_cells_layout + h2svd_expand_cells are the bulk form of these loops)

/// `GateChip::inner_product` of one row given its running sums: `[0, a_0, v_0, s_0, a_1, v_1, s_1, ...]`, gate at 3*j.
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

/// `range_check(x, n*lb)` given its witnesses `l0, l1, s1, l2, s2, ...`:
/// `[l0, l1, 2^lb, s1, l2, 2^2lb, s2, ...]`, gates at 0, 3, ...; every limb goes to the lookup table.
fn assign_range_check<F: BigPrimeField>(
    ctx: &mut Context<F>, range: &RangeChip<F>, x: AssignedValue<F>, n: usize, lb: usize, w: &[h2svd_fr],
) -> usize {
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

/// `check_big_less_than_safe(x, bound)`: range_check(x) | chk, xp | range_check(chk).
fn assign_cbls<F: BigPrimeField>(
    ctx: &mut Context<F>, range: &RangeChip<F>, x: AssignedValue<F>, bound: F, n: usize, lb: usize, w: &[h2svd_fr],
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

/// The cells of one `signed_div_scale(a)` from its W witnesses; returns the quotient cell.
fn assign_signed_div_scale<F: BigPrimeField>(
    ctx: &mut Context<F>, range: &RangeChip<F>, a: AssignedValue<F>, p: usize, lb: usize, s: usize, a_bits: usize,
    w: &[h2svd_fr],
) -> AssignedValue<F> {
    let pow = range.gate().pow_of_two();
    let (n_d, n_r) = ((a_bits - p + 1 + lb - 1) / lb, (p + 1 + lb - 1) / lb);
    let a_shift = ctx.assign_region_last([Existing(a), Constant(pow[s]), Constant(F::ONE), Witness(from_wire::<F>(&w[0]))], [0]);
    ctx.assign_region([Witness(from_wire::<F>(&w[1])), Constant(pow[p]), Witness(from_wire::<F>(&w[2])), Existing(a_shift)], [0]);
    let (rem, div) = (ctx.get(-4), ctx.get(-2));
    let mut used = 3;
    used += assign_cbls(ctx, range, div, pow[a_bits - p] + F::ONE, n_d, lb, &w[used..]);
    used += assign_cbls(ctx, range, rem, pow[p], n_r, lb, &w[used..]);
    ctx.assign_region([Witness(from_wire::<F>(&w[used])), Constant(pow[s - p]), Constant(F::ONE), Existing(div)], [0]);
    ctx.get(-4)
}

struct RescaleParams {
    p: usize,
    lb: usize,
    s: usize,
    a: usize,
    w: usize,
}
fn rescale_params(precision_bits: u32, lookup_bits: usize) -> RescaleParams {
    let p = precision_bits as usize;
    let (s, a) = (SHIFT_BITS.load(Ordering::Relaxed), A_NUM_BITS.load(Ordering::Relaxed));
    let s = if s < 0 { 3 * p } else { s as usize };
    let a = if a < 0 { 4 * p } else { a as usize };
    let w = unsafe { h2svd_rescale_witness_count(p as c_int, lookup_bits as c_int, s as c_int, a as c_int) };
    assert!(w > 0, "rescale parameters out of range");
    RescaleParams { p, lb: lookup_bits, s, a, w: w as usize }
}
/// One GPU call for `cells.len()` elements, then the cells of every `signed_div_scale` in order.
fn rescale_cells<F: BigPrimeField>(ctx: &mut Context<F>, range: &RangeChip<F>, precision_bits: u32, cells: &[AssignedValue<F>])
    -> Vec<AssignedValue<F>> {
    let rp = rescale_params(precision_bits, range.lookup_bits());
    let fc = gather1(cells);
    let mut q = vec![h2svd_fr::default(); cells.len()];
    let mut wit = vec![h2svd_fr::default(); cells.len() * rp.w];
    check(unsafe {
        h2svd_rescale_witness(gpu(), fc.as_ptr(), cells.len(), rp.p as c_int, rp.lb as c_int, rp.s as c_int, rp.a as c_int,
                              q.as_mut_ptr(), wit.as_mut_ptr())
    });
    cells.iter().enumerate()
        .map(|(e, c)| assign_signed_div_scale(ctx, range, *c, rp.p, rp.lb, rp.s, rp.a, &wit[e * rp.w..(e + 1) * rp.w]))
        .collect()
}

// ---------------------------------------------------------------------------------------------------------------------
// src/matrix/mod.rs:21-216

#[derive(Clone)]
pub struct ZkVector<F: BigPrimeField, const PRECISION_BITS: u32> {
    pub v: Vec<AssignedValue<F>>,
}

impl<F: BigPrimeField, const PRECISION_BITS: u32> ZkVector<F, PRECISION_BITS> {
    /// :29-40 -- quantization of the whole vector in one GPU call, one bulk `assign_witnesses`
    pub fn new(ctx: &mut Context<F>, _fpchip: &FixedPointChip041<F, PRECISION_BITS>, v: &Vec<f64>) -> Self {
        let mut q = vec![h2svd_fr::default(); v.len()];
        check(unsafe { h2svd_quantize(gpu(), v.as_ptr(), v.len(), PRECISION_BITS as c_int, q.as_mut_ptr()) });
        return Self { v: ctx.assign_witnesses(q.iter().map(from_wire::<F>)) };
    }

    /// :43
    pub fn size(&self) -> usize {
        return self.v.len();
    }

    /// :50-57 (host-only helper of the reference, unchanged)
    pub fn dequantize(&self, fpchip: &FixedPointChip041<F, PRECISION_BITS>) -> Vec<f64> {
        self.v.iter().map(|e| fpchip.dequantization(*e.value())).collect()
    }

    /// :79-106 -- running sums of `gate.inner_product(u = x, v = self)` from the GPU, then one `signed_div_scale`
    pub fn inner_product(
        &self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, x: &Vec<AssignedValue<F>>,
    ) -> AssignedValue<F> {
        assert!(self.size() == x.len()); // :86
        let (fx, fs) = (gather1(x), gather1(&self.v));
        let mut prefix = vec![h2svd_fr::default(); x.len()];
        check(unsafe { h2svd_zkvec_inner_prefix(gpu(), fx.as_ptr(), fs.as_ptr(), 1, x.len(), prefix.as_mut_ptr()) });
        let res_s = assign_inner_product_row(ctx, x, &self.v, &prefix); // :100
        return rescale_cells(ctx, fpchip.range_gate(), PRECISION_BITS, &[res_s])[0]; // :104
    }

    /// :111-117
    pub fn _norm_square(&self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>) -> AssignedValue<F> {
        return self.inner_product(ctx, fpchip, &self.v);
    }

    /// :124-131 -- `qsqrt` stays the chip's (third-party constraint layout)
    pub fn norm(&self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>) -> AssignedValue<F> {
        let norm_sq = self._norm_square(ctx, fpchip);
        return fpchip.qsqrt(ctx, norm_sq);
    }

    /// :136-149 -- all `qsub` values in one GPU call; `gate.sub` cells `[a - b, b, 1, a]`
    pub fn _dist_square(
        &self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, x: &Vec<AssignedValue<F>>,
    ) -> AssignedValue<F> {
        assert_eq!(self.size(), x.len()); // :142
        let (fs, fx) = (gather1(&self.v), gather1(x));
        let mut d = vec![h2svd_fr::default(); x.len()];
        check(unsafe { h2svd_zkvec_sub(gpu(), fs.as_ptr(), fx.as_ptr(), x.len(), d.as_mut_ptr()) });
        let mut diff: Vec<AssignedValue<F>> = Vec::with_capacity(x.len());
        for i in 0..x.len() {
            ctx.assign_region([Witness(from_wire::<F>(&d[i])), Existing(x[i]), Constant(F::ONE), Existing(self.v[i])], [0]);
            diff.push(ctx.get(-4)); // GateChip::sub returns the first cell
        }
        let diff = Self { v: diff };
        return diff._norm_square(ctx, fpchip);
    }

    /// :156-164
    pub fn dist(
        &self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, x: &Vec<AssignedValue<F>>,
    ) -> AssignedValue<F> {
        let dist_sq = self._dist_square(ctx, fpchip, x);
        return fpchip.qsqrt(ctx, dist_sq);
    }

    /// :169-182 -- `a.v`: ONE mat-vec call for all running sums and ONE rescale call for all rows, cells in the reference's
    /// order (row by row: inner-product cells, then that row's signed_div_scale cells)
    pub fn mul(&self, ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, a: &ZkMatrix<F, PRECISION_BITS>) -> Self {
        assert_eq!(a.num_col, self.size()); // :175
        let (rows, len) = (a.num_rows, a.num_col);
        let (fa, fv) = (gather(&a.matrix), gather1(&self.v));
        let mut prefix = vec![h2svd_fr::default(); rows * len];
        // inner_product(u = row, v = self): the products are row[t] * self[t]
        check(unsafe { h2svd_mat_vec_prefix(gpu(), fa.as_ptr(), fv.as_ptr(), rows, len, prefix.as_mut_ptr()) });
        let rp = rescale_params(PRECISION_BITS, fpchip.range_gate().lookup_bits());
        let totals: Vec<h2svd_fr> = (0..rows).map(|i| prefix[i * len + len - 1]).collect();
        let mut q = vec![h2svd_fr::default(); rows];
        let mut wit = vec![h2svd_fr::default(); rows * rp.w];
        check(unsafe {
            h2svd_rescale_witness(gpu(), totals.as_ptr(), rows, rp.p as c_int, rp.lb as c_int, rp.s as c_int, rp.a as c_int,
                                  q.as_mut_ptr(), wit.as_mut_ptr())
        });
        let mut y: Vec<AssignedValue<F>> = Vec::with_capacity(rows);
        for (i, row) in a.matrix.iter().enumerate() {
            let res_s = assign_inner_product_row(ctx, row, &self.v, &prefix[i * len..(i + 1) * len]);
            y.push(assign_signed_div_scale(ctx, fpchip.range_gate(), res_s, rp.p, rp.lb, rp.s, rp.a, &wit[i * rp.w..(i + 1) * rp.w]));
        }
        return Self { v: y };
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// src/matrix/mod.rs:219-419

#[derive(Clone)]
pub struct ZkMatrix<F: BigPrimeField, const PRECISION_BITS: u32> {
    pub matrix: Vec<Vec<AssignedValue<F>>>,
    pub num_rows: usize,
    pub num_col: usize,
}

impl<F: BigPrimeField, const PRECISION_BITS: u32> ZkMatrix<F, PRECISION_BITS> {
    /// :230-252
    pub fn new(ctx: &mut Context<F>, _fpchip: &FixedPointChip041<F, PRECISION_BITS>, matrix: &Vec<Vec<f64>>) -> Self {
        let num_rows = matrix.len();
        let num_col = matrix[0].len();
        for row in matrix {
            assert!(row.len() == num_col); // :239
        }
        let flat: Vec<f64> = matrix.iter().flatten().copied().collect();
        let mut q = vec![h2svd_fr::default(); flat.len()];
        check(unsafe { h2svd_quantize(gpu(), flat.as_ptr(), flat.len(), PRECISION_BITS as c_int, q.as_mut_ptr()) });
        let zkmatrix = q.chunks(num_col).map(|row| ctx.assign_witnesses(row.iter().map(from_wire::<F>))).collect();
        return Self { matrix: zkmatrix, num_rows, num_col };
    }

    /// :299-342
    pub fn verify_mul(
        ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, a: &Self, b: &Self,
        c_s: &Vec<Vec<AssignedValue<F>>>, init_rand: &AssignedValue<F>,
    ) {
        assert_eq!(a.num_col, b.num_rows); // :307
        assert_eq!(c_s.len(), a.num_rows); // :308
        assert_eq!(c_s[0].len(), b.num_col); // :309
        assert!(c_s[0].len() >= 1); // :310
        let (n, k, m) = (a.num_rows, a.num_col, b.num_col);
        let gate = fpchip.gate();
        let (fa, fb, fc) = (gather(&a.matrix), gather(&b.matrix), gather(c_s));
        let g = to_wire(init_rand.value());
        let z = h2svd_fr::default();
        let (mut pw, mut pcv, mut pbv, mut pabv) = (vec![z; m], vec![z; n * m], vec![z; k * m], vec![z; n * k]);
        let (mut diff, mut isz, mut inv) = (vec![z; n], vec![z; n], vec![z; n]);
        check(unsafe {
            h2svd_freivalds_witness(gpu(), fa.as_ptr(), fb.as_ptr(), fc.as_ptr(), &g, n, k, m, pw.as_mut_ptr(),
                                    pcv.as_mut_ptr(), pbv.as_mut_ptr(), pabv.as_mut_ptr(), diff.as_mut_ptr(),
                                    isz.as_mut_ptr(), inv.as_mut_ptr())
        });
        // :318-326  one, then v_i = v_{i-1} * gamma  (gate.mul: [0, v_{i-1}, gamma, v_i], gate at 0)
        let one = ctx.load_witness(from_wire::<F>(&pw[0]));
        gate.assert_is_const(ctx, &one, &F::ONE);
        let mut v = vec![one];
        for i in 1..m {
            let cells = [Constant(F::ZERO), Existing(v[i - 1]), Existing(*init_rand), Witness(from_wire::<F>(&pw[i]))];
            v.push(ctx.assign_region_last(cells, [0]));
        }
        // :335-337
        let cs_v: Vec<_> = (0..n).map(|i| assign_inner_product_row(ctx, &c_s[i], &v, &pcv[i * m..(i + 1) * m])).collect();
        let b_v: Vec<_> = (0..k).map(|i| assign_inner_product_row(ctx, &b.matrix[i], &v, &pbv[i * m..(i + 1) * m])).collect();
        let ab_v: Vec<_> = (0..n).map(|i| assign_inner_product_row(ctx, &a.matrix[i], &b_v, &pabv[i * k..(i + 1) * k])).collect();
        // :339-341  gate.is_equal(cs_v, ab_v) = GateChip::sub ([a - b, b, 1, a], which RETURNS ITS FIRST CELL) + is_zero
        for i in 0..n {
            ctx.assign_region(
                [Witness(from_wire::<F>(&diff[i])), Existing(ab_v[i]), Constant(F::ONE), Existing(cs_v[i])], [0]);
            let d = ctx.get(-4); // the diff witness -- NOT the last cell of the region (that is cs_v[i])
            let (zf, invf): (F, F) = (from_wire(&isz[i]), from_wire(&inv[i]));
            ctx.assign_region(
                [Witness(zf), Existing(d), Witness(invf), Constant(F::ONE), Constant(F::ZERO), Existing(d),
                 Witness(zf), Constant(F::ZERO)], [0, 4]);
        }
    }

    /// :354-375 -- ONE GPU call for the whole matrix, then the cells of `signed_div_scale` per element in the reference's order
    pub fn rescale_matrix(
        ctx: &mut Context<F>, fpchip: &FixedPointChip041<F, PRECISION_BITS>, c_s: &Vec<Vec<AssignedValue<F>>>,
    ) -> Self {
        let num_rows = c_s.len();
        let num_col = c_s[0].len();
        let flat: Vec<AssignedValue<F>> = c_s.iter().flatten().copied().collect();
        let q = rescale_cells(ctx, fpchip.range_gate(), PRECISION_BITS, &flat);
        let c = q.chunks(num_col).map(|r| r.to_vec()).collect();
        return Self { matrix: c, num_rows, num_col };
    }

    /// :408-419 (layout only, unchanged)
    pub fn transpose_matrix(a: &Self) -> Self {
        let mut a_trans: Vec<Vec<AssignedValue<F>>> = Vec::new();
        for i in 0..a.num_col {
            a_trans.push((0..a.num_rows).map(|j| a.matrix[j][i]).collect());
        }
        return Self { matrix: a_trans, num_rows: a.num_col, num_col: a.num_rows };
    }
}

/// :510-537 -- values only
pub fn field_mat_mul<F: BigPrimeField>(a: &Vec<Vec<AssignedValue<F>>>, b: &Vec<Vec<AssignedValue<F>>>) -> Vec<Vec<F>> {
    assert_eq!(a[0].len(), b.len()); // :515
    let (n, k, m) = (a.len(), b.len(), b[0].len());
    let (fa, fb) = (gather(a), gather(b));
    let mut c = vec![h2svd_fr::default(); n * m];
    check(unsafe { h2svd_fr_matmul(gpu(), fa.as_ptr(), fb.as_ptr(), c.as_mut_ptr(), n, k, m, 0) });
    c.chunks(m).map(|row| row.iter().map(from_wire::<F>).collect()).collect()
}

/// :546-568 -- N*M unconstrained witnesses, row-major; `assign_witnesses` is the bulk form of the reference's
/// `load_witness` loop (:558-565)
pub fn honest_prover_mat_mul<F: BigPrimeField>(
    ctx: &mut Context<F>, a: &Vec<Vec<AssignedValue<F>>>, b: &Vec<Vec<AssignedValue<F>>>,
) -> Vec<Vec<AssignedValue<F>>> {
    field_mat_mul(a, b).into_iter().map(|row| ctx.assign_witnesses(row)).collect()
}

/// :574-599
pub fn field_mat_vec_mul<F: BigPrimeField>(
    ctx: &mut Context<F>, _gate: &GateChip<F>, a: &Vec<Vec<AssignedValue<F>>>, v: &Vec<AssignedValue<F>>,
) -> Vec<AssignedValue<F>> {
    assert_eq!(a[0].len(), v.len()); // :580
    let (rows, len) = (a.len(), v.len());
    let (fa, fv) = (gather(a), gather1(v));
    let mut prefix = vec![h2svd_fr::default(); rows * len];
    check(unsafe { h2svd_mat_vec_prefix(gpu(), fa.as_ptr(), fv.as_ptr(), rows, len, prefix.as_mut_ptr()) });
    (0..rows).map(|i| assign_inner_product_row(ctx, &a[i], v, &prefix[i * len..(i + 1) * len])).collect()
}
