// Carry-chain (PTX) field primitives for the witness kernels outside the mat-mul inner loop:
// Montgomery multiplication, single-limb Montgomery step, modular add/sub.
//
// Representation used inside a multiplication: the running value V is held as two base-2^64 numbers
//   V = E + O * 2^32,   E = sum_p E[p] * 2^(64p)   (64-bit columns at EVEN 32-bit limb positions)
//                       O = sum_p O[p] * 2^(64p)   (64-bit columns at ODD  32-bit limb positions)
// so that every 32x32->64 product lands on an aligned register pair and a whole row
// {a0,a1,a2,a3} * b is ONE carry chain of four fused `IMAD.WIDE.U32.X` (mad.lo.cc + madc.hi.cc on a
// register pair) -- the same instruction shape the mat-mul inner loop uses (fr_acc.cuh).
//
// mont_mul_fast runs CIOS two multiplier limbs per step and then drops 64 bits (one whole column
// of E and of O), so columns never change parity; the single 32-bit value and the single carry bit
// that fall off O[0] are folded into the first chain of the next step.  Cost: 132 IMAD.WIDE + 8 IMAD
// per field multiplication (vs ~2x that plus 64-bit emulation overhead for the portable fr::mont_mul).
//
// Every asm block is self-contained with respect to the carry flag, and every block has a host
// fallback that restates it with unsigned __int128, so the ALGORITHM (column bookkeeping, carry
// folding, bounds) is unit-tested on the CPU build box (tests/test_fr_host.py); the GPU parity tests
// cover the PTX transcription.
#pragma once
#include "fr.cuh"

namespace fr {

FR_HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
FR_HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }

// d[0..3] += {a0,a1,a2,a3} * b  (one carry chain over four 64-bit columns);  d[4] += carry out.
// PRECONDITION: d[4] < 2^32 - 1 (it only ever collects carries in the callers below), so the carry is added to
// its low word alone.
FR_HD void row4(uint64_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3, l4, h4;\n\t"
        "mov.b64 {l0, h0}, %0;\n\t"
        "mov.b64 {l1, h1}, %1;\n\t"
        "mov.b64 {l2, h2}, %2;\n\t"
        "mov.b64 {l3, h3}, %3;\n\t"
        "mov.b64 {l4, h4}, %4;\n\t"
        "mad.lo.cc.u32   l0, %5, %9, l0;\n\t"
        "madc.hi.cc.u32  h0, %5, %9, h0;\n\t"
        "madc.lo.cc.u32  l1, %6, %9, l1;\n\t"
        "madc.hi.cc.u32  h1, %6, %9, h1;\n\t"
        "madc.lo.cc.u32  l2, %7, %9, l2;\n\t"
        "madc.hi.cc.u32  h2, %7, %9, h2;\n\t"
        "madc.lo.cc.u32  l3, %8, %9, l3;\n\t"
        "madc.hi.cc.u32  h3, %8, %9, h3;\n\t"
        "addc.u32        l4, l4, 0;\n\t"
        "mov.b64 %0, {l0, h0};\n\t"
        "mov.b64 %1, {l1, h1};\n\t"
        "mov.b64 %2, {l2, h2};\n\t"
        "mov.b64 %3, {l3, h3};\n\t"
        "mov.b64 %4, {l4, h4};\n\t"
        "}"
        : "+l"(d[0]), "+l"(d[1]), "+l"(d[2]), "+l"(d[3]), "+l"(d[4])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    const uint32_t a[4] = {a0, a1, a2, a3};
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a[i] * b) + d[i] + carry;
        d[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    d[4] += carry;
#endif
}

// Same as row4, but the first column additionally absorbs the 32-bit value x and the carry bit of
// (ca + cb) -- what fell off the bottom of O when the previous step dropped 64 bits.
// a0*b + x <= (2^32-1)^2 + 2^32-1 < 2^64, so the fold cannot overflow.
FR_HD void row4_fold(uint64_t* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b,
                     uint32_t x, uint32_t ca, uint32_t cb) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3, l4, h4, tl, th, junk;\n\t"
        ".reg .u64 xx, tt;\n\t"
        "mov.b64 {l0, h0}, %0;\n\t"
        "mov.b64 {l1, h1}, %1;\n\t"
        "mov.b64 {l2, h2}, %2;\n\t"
        "mov.b64 {l3, h3}, %3;\n\t"
        "mov.b64 {l4, h4}, %4;\n\t"
        "cvt.u64.u32 xx, %10;\n\t"
        "mad.wide.u32 tt, %5, %9, xx;\n\t"
        "mov.b64 {tl, th}, tt;\n\t"
        "add.cc.u32      junk, %11, %12;\n\t"
        "addc.cc.u32     l0, l0, tl;\n\t"
        "addc.cc.u32     h0, h0, th;\n\t"
        "madc.lo.cc.u32  l1, %6, %9, l1;\n\t"
        "madc.hi.cc.u32  h1, %6, %9, h1;\n\t"
        "madc.lo.cc.u32  l2, %7, %9, l2;\n\t"
        "madc.hi.cc.u32  h2, %7, %9, h2;\n\t"
        "madc.lo.cc.u32  l3, %8, %9, l3;\n\t"
        "madc.hi.cc.u32  h3, %8, %9, h3;\n\t"
        "addc.u32        l4, l4, 0;\n\t"
        "mov.b64 %0, {l0, h0};\n\t"
        "mov.b64 %1, {l1, h1};\n\t"
        "mov.b64 %2, {l2, h2};\n\t"
        "mov.b64 %3, {l3, h3};\n\t"
        "mov.b64 %4, {l4, h4};\n\t"
        "}"
        : "+l"(d[0]), "+l"(d[1]), "+l"(d[2]), "+l"(d[3]), "+l"(d[4])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b), "r"(x), "r"(ca), "r"(cb));
#else
    const uint32_t a[4] = {a0, a1, a2, a3};
    uint64_t carry = ((uint64_t)ca + cb) >> 32;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a[i] * b) + d[i] + carry + (i == 0 ? x : 0u);
        d[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    d[4] += carry;
#endif
}

// a (8 limbs, < 2r) -> a - r if a >= r
FR_HD void cond_sub_r(uint32_t* a) {
#if defined(__CUDA_ARCH__)
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, bw;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(modulus(0)),
          "r"(modulus(1)), "r"(modulus(2)), "r"(modulus(3)), "r"(modulus(4)), "r"(modulus(5)), "r"(modulus(6)),
          "r"(modulus(7)));
    const bool keep = bw != 0;  // borrow: a < r
    a[0] = keep ? a[0] : t0;
    a[1] = keep ? a[1] : t1;
    a[2] = keep ? a[2] : t2;
    a[3] = keep ? a[3] : t3;
    a[4] = keep ? a[4] : t4;
    a[5] = keep ? a[5] : t5;
    a[6] = keep ? a[6] : t6;
    a[7] = keep ? a[7] : t7;
#else
    cond_sub_mod(a);
#endif
}

// (a + b) mod r for canonical a, b
FR_HD Fr add_fast(const Fr& a, const Fr& b) {
    Fr o;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=r"(o.l[0]), "=r"(o.l[1]), "=r"(o.l[2]), "=r"(o.l[3]), "=r"(o.l[4]), "=r"(o.l[5]), "=r"(o.l[6]),
          "=r"(o.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
    add_n<8>(o.l, a.l, b.l);
#endif
    cond_sub_r(o.l);  // a + b < 2r < 2^255: no carry out of limb 7
    return o;
}

// (a - b) mod r for canonical a, b
FR_HD Fr sub_fast(const Fr& a, const Fr& b) {
#if defined(__CUDA_ARCH__)
    Fr o;
    uint32_t bw;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(o.l[0]), "=r"(o.l[1]), "=r"(o.l[2]), "=r"(o.l[3]), "=r"(o.l[4]), "=r"(o.l[5]), "=r"(o.l[6]),
          "=r"(o.l[7]), "=r"(bw)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // bw = 0xffffffff when a < b: add r back (masked)
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(o.l[0]), "+r"(o.l[1]), "+r"(o.l[2]), "+r"(o.l[3]), "+r"(o.l[4]), "+r"(o.l[5]), "+r"(o.l[6]),
          "+r"(o.l[7])
        : "r"(modulus(0) & bw), "r"(modulus(1) & bw), "r"(modulus(2) & bw), "r"(modulus(3) & bw),
          "r"(modulus(4) & bw), "r"(modulus(5) & bw), "r"(modulus(6) & bw), "r"(modulus(7) & bw));
    return o;
#else
    return sub(a, b);
#endif
}

// out[j] = x[j] + y[j] (8 limbs, one chain) + carry bit of (ca + cb) into limb 0; the sum must fit.
FR_HD void add8_carry_in(uint32_t* o, const uint32_t* x, const uint32_t* y, uint32_t ca, uint32_t cb) {
#if defined(__CUDA_ARCH__)
    uint32_t junk;
    asm("add.cc.u32  %8, %25, %26;\n\t"
        "addc.cc.u32 %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.u32    %7, %16, %24;"
        : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]),
          "=r"(junk)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]),
          "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(ca), "r"(cb));
#else
    uint64_t c = ((uint64_t)ca + cb) >> 32;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x[i] + y[i];
        o[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// Montgomery product a * b * 2^-256 mod r, canonical; a, b canonical (< r).
FR_HD Fr mont_mul_fast(const Fr& a, const Fr& b) {
    uint64_t E[6] = {0, 0, 0, 0, 0, 0};  // columns at limbs (0,1) (2,3) (4,5) (6,7) (8,9) (10,11)
    uint64_t O[5] = {0, 0, 0, 0, 0};     // columns at limbs (1,2) (3,4) (5,6) (7,8) (9,10)
    uint32_t x = 0, ca = 0, cb = 0;      // value at limb 0 and carry bit left over from the previous step
    const uint32_t r0 = modulus(0), r1 = modulus(1), r2 = modulus(2), r3 = modulus(3), r4 = modulus(4),
                   r5 = modulus(5), r6 = modulus(6), r7 = modulus(7);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        // multiplier limb b[i], weight 2^0 in the current frame
        row4_fold(&E[0], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i], x, ca, cb);
        row4(&O[0], a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
        const uint32_t m0 = lo32(E[0]) * INV32;  // limb 0 of V is the low half of E[0] only
        row4(&E[0], r0, r2, r4, r6, m0);          // limb 0 becomes 0, its carry sits inside E[0]
        row4(&O[0], r1, r3, r5, r7, m0);
        // multiplier limb b[i+1], weight 2^32: even a-limbs land on odd columns and vice versa
        row4(&O[0], a.l[0], a.l[2], a.l[4], a.l[6], b.l[i + 1]);
        row4(&E[1], a.l[1], a.l[3], a.l[5], a.l[7], b.l[i + 1]);
        const uint32_t m1 = (hi32(E[0]) + lo32(O[0])) * INV32;  // limb 1 of V
        row4(&O[0], r0, r2, r4, r6, m1);
        row4(&E[1], r1, r3, r5, r7, m1);
        // limbs 0 and 1 are now zero: hi32(E[0]) + lo32(O[0]) is 0 or 2^32 (one carry bit into limb 2);
        // hi32(O[0]) is the rest of limb 2.  Drop 64 bits.
        ca = hi32(E[0]);
        cb = lo32(O[0]);
        x = hi32(O[0]);
#pragma unroll
        for (int p = 0; p < 5; p++) E[p] = E[p + 1];
        E[5] = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) O[p] = O[p + 1];
        O[4] = 0;
    }
    // V = E + O*2^32 + x + carry(ca + cb) < 2r
    const uint32_t ev[8] = {lo32(E[0]), hi32(E[0]), lo32(E[1]), hi32(E[1]), lo32(E[2]), hi32(E[2]), lo32(E[3]), hi32(E[3])};
    const uint32_t ov[8] = {x, lo32(O[0]), hi32(O[0]), lo32(O[1]), hi32(O[1]), lo32(O[2]), hi32(O[2]), lo32(O[3])};
    Fr o;
    add8_carry_in(o.l, ev, ov, ca, cb);
    cond_sub_r(o.l);
    return o;
}

// Two independent Montgomery products with their carry chains INTERLEAVED in program order: one product is a single long
// dependent chain per column set (every row4 waits for the previous one through the carry predicate and the accumulator
// registers), so a warp issues at most ~2 IMAD.WIDE per chain latency; two products give the scheduler four independent
// chains.  Same arithmetic as mont_mul_fast, element for element.
FR_HD void mont_mul_fast_x2(const Fr& a0, const Fr& b0, const Fr& a1, const Fr& b1, Fr& o0, Fr& o1) {
    uint64_t E0[6] = {0, 0, 0, 0, 0, 0}, O0[5] = {0, 0, 0, 0, 0};
    uint64_t E1[6] = {0, 0, 0, 0, 0, 0}, O1[5] = {0, 0, 0, 0, 0};
    uint32_t x0 = 0, ca0 = 0, cb0 = 0, x1 = 0, ca1 = 0, cb1 = 0;
    const uint32_t r0 = modulus(0), r1 = modulus(1), r2 = modulus(2), r3 = modulus(3), r4 = modulus(4),
                   r5 = modulus(5), r6 = modulus(6), r7 = modulus(7);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        row4_fold(&E0[0], a0.l[0], a0.l[2], a0.l[4], a0.l[6], b0.l[i], x0, ca0, cb0);
        row4_fold(&E1[0], a1.l[0], a1.l[2], a1.l[4], a1.l[6], b1.l[i], x1, ca1, cb1);
        row4(&O0[0], a0.l[1], a0.l[3], a0.l[5], a0.l[7], b0.l[i]);
        row4(&O1[0], a1.l[1], a1.l[3], a1.l[5], a1.l[7], b1.l[i]);
        const uint32_t m00 = lo32(E0[0]) * INV32, m01 = lo32(E1[0]) * INV32;
        row4(&E0[0], r0, r2, r4, r6, m00);
        row4(&E1[0], r0, r2, r4, r6, m01);
        row4(&O0[0], r1, r3, r5, r7, m00);
        row4(&O1[0], r1, r3, r5, r7, m01);
        row4(&O0[0], a0.l[0], a0.l[2], a0.l[4], a0.l[6], b0.l[i + 1]);
        row4(&O1[0], a1.l[0], a1.l[2], a1.l[4], a1.l[6], b1.l[i + 1]);
        row4(&E0[1], a0.l[1], a0.l[3], a0.l[5], a0.l[7], b0.l[i + 1]);
        row4(&E1[1], a1.l[1], a1.l[3], a1.l[5], a1.l[7], b1.l[i + 1]);
        const uint32_t m10 = (hi32(E0[0]) + lo32(O0[0])) * INV32, m11 = (hi32(E1[0]) + lo32(O1[0])) * INV32;
        row4(&O0[0], r0, r2, r4, r6, m10);
        row4(&O1[0], r0, r2, r4, r6, m11);
        row4(&E0[1], r1, r3, r5, r7, m10);
        row4(&E1[1], r1, r3, r5, r7, m11);
        ca0 = hi32(E0[0]); cb0 = lo32(O0[0]); x0 = hi32(O0[0]);
        ca1 = hi32(E1[0]); cb1 = lo32(O1[0]); x1 = hi32(O1[0]);
#pragma unroll
        for (int p = 0; p < 5; p++) {
            E0[p] = E0[p + 1];
            E1[p] = E1[p + 1];
        }
        E0[5] = 0;
        E1[5] = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            O0[p] = O0[p + 1];
            O1[p] = O1[p + 1];
        }
        O0[4] = 0;
        O1[4] = 0;
    }
    {
        const uint32_t ev[8] = {lo32(E0[0]), hi32(E0[0]), lo32(E0[1]), hi32(E0[1]), lo32(E0[2]), hi32(E0[2]), lo32(E0[3]), hi32(E0[3])};
        const uint32_t ov[8] = {x0, lo32(O0[0]), hi32(O0[0]), lo32(O0[1]), hi32(O0[1]), lo32(O0[2]), hi32(O0[2]), lo32(O0[3])};
        add8_carry_in(o0.l, ev, ov, ca0, cb0);
        cond_sub_r(o0.l);
    }
    {
        const uint32_t ev[8] = {lo32(E1[0]), hi32(E1[0]), lo32(E1[1]), hi32(E1[1]), lo32(E1[2]), hi32(E1[2]), lo32(E1[3]), hi32(E1[3])};
        const uint32_t ov[8] = {x1, lo32(O1[0]), hi32(O1[0]), lo32(O1[1]), hi32(O1[1]), lo32(O1[2]), hi32(O1[2]), lo32(O1[3])};
        add8_carry_in(o1.l, ev, ov, ca1, cb1);
        cond_sub_r(o1.l);
    }
}

// Single-limb Montgomery step: l * c * 2^-32 mod r, canonical (l any u32, c canonical).
// With c = x * 2^256 * 2^32 mod r this yields the Montgomery form of l * x in 16 IMAD.WIDE + 1 IMAD.
FR_HD Fr mont_mul_small(uint32_t l, const Fr& c) {
    uint64_t E[5], O[5];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        E[p] = (uint64_t)c.l[2 * p] * l;
        O[p] = (uint64_t)c.l[2 * p + 1] * l;
    }
    E[4] = 0;
    O[4] = 0;
    const uint32_t m = lo32(E[0]) * INV32;
    row4(E, modulus(0), modulus(2), modulus(4), modulus(6), m);
    row4(O, modulus(1), modulus(3), modulus(5), modulus(7), m);
    // (E + O*2^32) >> 32, < 2r
    const uint32_t ev[8] = {hi32(E[0]), lo32(E[1]), hi32(E[1]), lo32(E[2]), hi32(E[2]), lo32(E[3]), hi32(E[3]), lo32(E[4])};
    const uint32_t ov[8] = {lo32(O[0]), hi32(O[0]), lo32(O[1]), hi32(O[1]), lo32(O[2]), hi32(O[2]), lo32(O[3]), hi32(O[3])};
    Fr o;
    add8_carry_in(o.l, ev, ov, 0u, 0u);
    // o < l * c / 2^32 + r: for the small l of a limb decomposition (l < 2^lb) o exceeds r about once in 2^(32 - lb) calls,
    // and o >= r needs o's top limb >= r's: a rarely-taken branch replaces the 17-instruction compare-and-select
    if (o.l[7] >= modulus(7)) cond_sub_r(o.l);
    return o;
}

// ---- running sums of small multiples of fixed constants, every partial sum canonical --------------------------------
// s_i = sum_{j<=i} l_j * rho_j mod r for 32-bit l_j and fixed canonical rho_j (range-check running sums: rho_j = the
// Montgomery form of 2^(lb*j)), WITHOUT a Montgomery step and a compare-and-subtract per term:
//   U_i = sum l_j * rho_j            the unreduced integer (< 2^32 * r when sum l_j < 2^32): 8 wide products per term,
//   F_i = sum l_j * phi_j            phi_j = floor(rho_j * 2^64 / r): a 64-bit fixed-point estimate of U_i / r from below,
//   Q   = floor(F_i / 2^64)          = floor(U_i / r) or one less; it can be one less only when the fraction of F_i is
//                                    within sum l_j / 2^64 < 2^-32 of 1,
//   s_i = U_i - Q * r = U_i + Q * (2^256 - r) mod 2^256   (8 wide products; s_i < 2r < 2^256 so the wrap is exact)
// and a compare-and-subtract only in the rare case the top 20 fraction bits are all ones.  The partial sums are independent
// of each other given (U_i, F_i): no dependent chain through the canonical values.
struct SmallSum {
    uint64_t E[5], O[5];     // U = E + O * 2^32 (64-bit columns at even / odd limbs, as in mont_mul_fast)
    uint32_t f0, f1, f2;     // F = f0 + f1 * 2^32 + f2 * 2^64
};
FR_HD void small_sum_init(SmallSum& s) {
#pragma unroll
    for (int p = 0; p < 5; p++) s.E[p] = s.O[p] = 0;
    s.f0 = s.f1 = s.f2 = 0;
}
FR_HD void small_sum_add(SmallSum& s, uint32_t l, const Fr& rho, uint32_t phi_lo, uint32_t phi_hi) {
    row4(s.E, rho.l[0], rho.l[2], rho.l[4], rho.l[6], l);
    row4(s.O, rho.l[1], rho.l[3], rho.l[5], rho.l[7], l);
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;\n\t"
        "mad.lo.cc.u32  %1, %3, %5, %1;\n\t"
        "madc.hi.u32    %2, %3, %5, %2;"
        : "+r"(s.f0), "+r"(s.f1), "+r"(s.f2)
        : "r"(l), "r"(phi_lo), "r"(phi_hi));
#else
    const unsigned __int128 f = ((unsigned __int128)s.f2 << 64) | ((uint64_t)s.f1 << 32) | s.f0;
    const unsigned __int128 g = f + (unsigned __int128)l * (((uint64_t)phi_hi << 32) | phi_lo);
    s.f0 = (uint32_t)g;
    s.f1 = (uint32_t)(g >> 32);
    s.f2 = (uint32_t)(g >> 64);
#endif
}
FR_HD Fr small_sum_value(const SmallSum& s) {
    uint64_t E[5], O[5];
#pragma unroll
    for (int p = 0; p < 5; p++) {
        E[p] = s.E[p];
        O[p] = s.O[p];
    }
    // 2^256 - r: r is odd, so only limb 0 takes the +1 of the two's complement
    row4(E, 0u - modulus(0), ~modulus(2), ~modulus(4), ~modulus(6), s.f2);
    row4(O, ~modulus(1), ~modulus(3), ~modulus(5), ~modulus(7), s.f2);
    const uint32_t ev[8] = {lo32(E[0]), hi32(E[0]), lo32(E[1]), hi32(E[1]), lo32(E[2]), hi32(E[2]), lo32(E[3]), hi32(E[3])};
    const uint32_t ov[8] = {0u, lo32(O[0]), hi32(O[0]), lo32(O[1]), hi32(O[1]), lo32(O[2]), hi32(O[2]), lo32(O[3])};
    Fr o;
    add8_carry_in(o.l, ev, ov, 0u, 0u);   // mod 2^256
    if (s.f1 >= 0xFFFFF000u) cond_sub_r(o.l);
    return o;
}

FR_HD Fr to_mont_fast(const Fr& x) {
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = mont_r2(i);
    return mont_mul_fast(x, r2);
}
FR_HD Fr from_mont_fast(const Fr& a) {
    Fr one_int = zero();
    one_int.l[0] = 1u;
    return mont_mul_fast(a, one_int);
}

// Montgomery reduction alone: a * 2^-256 mod r, canonical (a canonical) -- the same value as from_mont_fast(a), without the
// 64 multiplier products a * 1 that mont_mul_fast would still issue (80 instead of 132 IMAD.WIDE).  V starts as a itself.
FR_HD Fr mont_reduce_fast(const Fr& a) {
    uint64_t E[6] = {(uint64_t)a.l[0] | ((uint64_t)a.l[1] << 32), (uint64_t)a.l[2] | ((uint64_t)a.l[3] << 32),
                     (uint64_t)a.l[4] | ((uint64_t)a.l[5] << 32), (uint64_t)a.l[6] | ((uint64_t)a.l[7] << 32), 0, 0};
    uint64_t O[5] = {0, 0, 0, 0, 0};
    uint32_t x = 0, ca = 0, cb = 0;
    const uint32_t r0 = modulus(0), r1 = modulus(1), r2 = modulus(2), r3 = modulus(3), r4 = modulus(4),
                   r5 = modulus(5), r6 = modulus(6), r7 = modulus(7);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        row4_fold(&E[0], 0u, 0u, 0u, 0u, 0u, x, ca, cb);  // fold what fell off O in the previous step (no multiplier row)
        const uint32_t m0 = lo32(E[0]) * INV32;
        row4(&E[0], r0, r2, r4, r6, m0);
        row4(&O[0], r1, r3, r5, r7, m0);
        const uint32_t m1 = (hi32(E[0]) + lo32(O[0])) * INV32;
        row4(&O[0], r0, r2, r4, r6, m1);
        row4(&E[1], r1, r3, r5, r7, m1);
        ca = hi32(E[0]);
        cb = lo32(O[0]);
        x = hi32(O[0]);
#pragma unroll
        for (int p = 0; p < 5; p++) E[p] = E[p + 1];
        E[5] = 0;
#pragma unroll
        for (int p = 0; p < 4; p++) O[p] = O[p + 1];
        O[4] = 0;
    }
    const uint32_t ev[8] = {lo32(E[0]), hi32(E[0]), lo32(E[1]), hi32(E[1]), lo32(E[2]), hi32(E[2]), lo32(E[3]), hi32(E[3])};
    const uint32_t ov[8] = {x, lo32(O[0]), hi32(O[0]), lo32(O[1]), hi32(O[1]), lo32(O[2]), hi32(O[2]), lo32(O[3])};
    Fr o;
    add8_carry_in(o.l, ev, ov, ca, cb);
    cond_sub_r(o.l);
    return o;
}

}  // namespace fr
