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
