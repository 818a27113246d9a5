// One-level Karatsuba for the Fr mat-mul inner loop: 48 instead of 64 IMAD.WIDE per multiply-add.
//
// A canonical operand x < r < 2^254 is split at bit 127:  x = lo + hi * 2^127,  lo, hi < 2^127, and the sum
// s = lo + hi < 2^128 still fits four 32-bit limbs -- no carry bit to drag through the inner loop (a split at
// bit 128 would make s a 129-bit number).  With P0 = sum_k a.lo*b.lo, P2 = sum_k a.hi*b.hi and
// P1 = sum_k a.s*b.s accumulated LAZILY (three independent 4x4-limb product sums, no reduction in the k loop)
//     sum_k a*b = P0 + (P1 - P0 - P2) * 2^127 + P2 * 2^254
// is formed once per C element in the epilogue (Karatsuba's recombination is linear, so it commutes with the
// sum over k) and then reduced exactly like the schoolbook accumulator (fr::reduce_wide_acc).  Exact integer
// arithmetic throughout: the result is the same canonical field element, bit for bit.
//
// The operands are pre-split once per matrix (O(N^2)) into the 48-byte KOp layout {lo, hi, s}.
// As in fr_acc.cuh every 32x32->64 product lands on an aligned 64-bit column of an even or an odd accumulator so
// that mad.lo.cc + madc.hi.cc fuse into one IMAD.WIDE.U32[.X]; a 4x4 product is 8 chains of two products.
#pragma once
#include "fr.cuh"

namespace fr {

struct alignas(16) KOp {   // 48 bytes: the three 128-bit pieces of one field element
    uint32_t lo[4], hi[4], s[4];
};

struct KAcc {              // lazy sum of 4x4-limb products (< 2^(256 + 32) for < 2^30 terms)
    uint64_t e[4];         // columns at limbs (0,1) (2,3) (4,5) (6,7)
    uint64_t o[3];         // columns at limbs (1,2) (3,4) (5,6)
    uint32_t ce3, ce4;     // carries out of even chains, weights 2^(32*6), 2^(32*8)
    uint32_t co2, co3;     // carries out of odd chains,  weights 2^(32*5), 2^(32*7)
};

FR_HD void kacc_clear(KAcc& w) {
#pragma unroll
    for (int i = 0; i < 4; i++) w.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) w.o[i] = 0;
    w.ce3 = w.ce4 = w.co2 = w.co3 = 0;
}

FR_HD KOp ksplit(const Fr& x) {
    KOp k;
    k.lo[0] = x.l[0];
    k.lo[1] = x.l[1];
    k.lo[2] = x.l[2];
    k.lo[3] = x.l[3] & 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 4; i++) k.hi[i] = (x.l[3 + i] >> 31) | (x.l[4 + i] << 1);
    // hi[3] = (x.l[6] >> 31) | (x.l[7] << 1): x < 2^254 so x.l[7] < 2^30 and hi < 2^127
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c += (uint64_t)k.lo[i] + k.hi[i];
        k.s[i] = (uint32_t)c;
        c >>= 32;
    }
    return k;  // c == 0: lo + hi < 2^128
}

// d[0..N) (N consecutive 64-bit columns) += a_t * b_t for t < N, one carry chain; cnt += carry out.
// Every product may have its own (a, b): a chain only needs its products to sit on consecutive columns.
template <int N>
FR_HD void kchain(uint64_t* d, uint32_t& cnt, const uint32_t* a, const uint32_t* b) {
#if defined(__CUDA_ARCH__)
    static_assert(N >= 1 && N <= 4, "chain length");
    if (N == 1) {
        asm("{\n\t"
            ".reg .u32 l0, h0;\n\t"
            "mov.b64 {l0, h0}, %0;\n\t"
            "mad.lo.cc.u32   l0, %2, %3, l0;\n\t"
            "madc.hi.cc.u32  h0, %2, %3, h0;\n\t"
            "addc.u32        %1, %1, 0;\n\t"
            "mov.b64 %0, {l0, h0};\n\t"
            "}"
            : "+l"(d[0]), "+r"(cnt)
            : "r"(a[0]), "r"(b[0]));
    } else if (N == 2) {
        asm("{\n\t"
            ".reg .u32 l0, h0, l1, h1;\n\t"
            "mov.b64 {l0, h0}, %0;\n\t"
            "mov.b64 {l1, h1}, %1;\n\t"
            "mad.lo.cc.u32   l0, %3, %5, l0;\n\t"
            "madc.hi.cc.u32  h0, %3, %5, h0;\n\t"
            "madc.lo.cc.u32  l1, %4, %6, l1;\n\t"
            "madc.hi.cc.u32  h1, %4, %6, h1;\n\t"
            "addc.u32        %2, %2, 0;\n\t"
            "mov.b64 %0, {l0, h0};\n\t"
            "mov.b64 %1, {l1, h1};\n\t"
            "}"
            : "+l"(d[0]), "+l"(d[1]), "+r"(cnt)
            : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]));
    } else if (N == 3) {
        asm("{\n\t"
            ".reg .u32 l0, h0, l1, h1, l2, h2;\n\t"
            "mov.b64 {l0, h0}, %0;\n\t"
            "mov.b64 {l1, h1}, %1;\n\t"
            "mov.b64 {l2, h2}, %2;\n\t"
            "mad.lo.cc.u32   l0, %4, %7, l0;\n\t"
            "madc.hi.cc.u32  h0, %4, %7, h0;\n\t"
            "madc.lo.cc.u32  l1, %5, %8, l1;\n\t"
            "madc.hi.cc.u32  h1, %5, %8, h1;\n\t"
            "madc.lo.cc.u32  l2, %6, %9, l2;\n\t"
            "madc.hi.cc.u32  h2, %6, %9, h2;\n\t"
            "addc.u32        %3, %3, 0;\n\t"
            "mov.b64 %0, {l0, h0};\n\t"
            "mov.b64 %1, {l1, h1};\n\t"
            "mov.b64 %2, {l2, h2};\n\t"
            "}"
            : "+l"(d[0]), "+l"(d[1]), "+l"(d[2]), "+r"(cnt)
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(b[0]), "r"(b[1]), "r"(b[2]));
    } else {
        asm("{\n\t"
            ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
            "mov.b64 {l0, h0}, %0;\n\t"
            "mov.b64 {l1, h1}, %1;\n\t"
            "mov.b64 {l2, h2}, %2;\n\t"
            "mov.b64 {l3, h3}, %3;\n\t"
            "mad.lo.cc.u32   l0, %5, %9,  l0;\n\t"
            "madc.hi.cc.u32  h0, %5, %9,  h0;\n\t"
            "madc.lo.cc.u32  l1, %6, %10, l1;\n\t"
            "madc.hi.cc.u32  h1, %6, %10, h1;\n\t"
            "madc.lo.cc.u32  l2, %7, %11, l2;\n\t"
            "madc.hi.cc.u32  h2, %7, %11, h2;\n\t"
            "madc.lo.cc.u32  l3, %8, %12, l3;\n\t"
            "madc.hi.cc.u32  h3, %8, %12, h3;\n\t"
            "addc.u32        %4, %4, 0;\n\t"
            "mov.b64 %0, {l0, h0};\n\t"
            "mov.b64 %1, {l1, h1};\n\t"
            "mov.b64 %2, {l2, h2};\n\t"
            "mov.b64 %3, {l3, h3};\n\t"
            "}"
            : "+l"(d[0]), "+l"(d[1]), "+l"(d[2]), "+l"(d[3]), "+r"(cnt)
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    }
#else
    uint32_t carry = 0;
    for (int i = 0; i < N; i++) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a[i] * b[i]) + d[i] + carry;
        d[i] = (uint64_t)t;
        carry = (uint32_t)(t >> 64);
    }
    cnt += carry;
#endif
}

// w += a * b   (a, b: 4 x u32 limbs each): 16 IMAD.WIDE in 7 carry chains.  Product a_i*b_j sits at limb i+j: even
// sums on the e columns, odd sums on the o columns.  The chains are chosen so that the carry-outs land on only
// four counters, three of which receive two carries each (ptxas folds those into one IADD3.X with two carry
// predicates): the diagonal a_i*b_i is one 4-chain, the rest are 2- and 3-chains plus two single products.
FR_HD void kmul_acc(KAcc& w, const uint32_t* a, const uint32_t* b) {
    {   // even: (0,0) (1,1) (2,2) (3,3) on columns 0..3 -> limb 8
        kchain<4>(&w.e[0], w.ce4, a, b);
    }
    {   // even: (2,0) (3,1) on columns 1, 2 -> limb 6
        const uint32_t x[2] = {a[2], a[3]}, y[2] = {b[0], b[1]};
        kchain<2>(&w.e[1], w.ce3, x, y);
    }
    {   // even: (0,2) (1,3) on columns 1, 2 -> limb 6
        const uint32_t x[2] = {a[0], a[1]}, y[2] = {b[2], b[3]};
        kchain<2>(&w.e[1], w.ce3, x, y);
    }
    {   // odd: (1,0) (3,0) (3,2) on columns 0, 1, 2 -> limb 7
        const uint32_t x[3] = {a[1], a[3], a[3]}, y[3] = {b[0], b[0], b[2]};
        kchain<3>(&w.o[0], w.co3, x, y);
    }
    {   // odd: (0,1) (2,1) (2,3) on columns 0, 1, 2 -> limb 7
        const uint32_t x[3] = {a[0], a[2], a[2]}, y[3] = {b[1], b[1], b[3]};
        kchain<3>(&w.o[0], w.co3, x, y);
    }
    {   // odd: (1,2) and (0,3) on column 1 -> limb 5
        kchain<1>(&w.o[1], w.co2, &a[1], &b[2]);
        kchain<1>(&w.o[1], w.co2, &a[0], &b[3]);
    }
}

// redundant form -> plain 10-limb integer
FR_HD void kacc_collapse(const KAcc& w, uint32_t* T) {
    uint64_t col[10];
#pragma unroll
    for (int i = 0; i < 10; i++) col[i] = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        col[2 * p] += (uint32_t)w.e[p];
        col[2 * p + 1] += (uint32_t)(w.e[p] >> 32);
    }
#pragma unroll
    for (int p = 0; p < 3; p++) {
        col[2 * p + 1] += (uint32_t)w.o[p];
        col[2 * p + 2] += (uint32_t)(w.o[p] >> 32);
    }
    col[6] += w.ce3;
    col[8] += w.ce4;
    col[5] += w.co2;
    col[7] += w.co3;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        c += col[i];
        T[i] = (uint32_t)c;
        c >>= 32;
    }
}

// T[0..18) += src[0..10) << shift   (shift = 32*ws + bs, all constants)
template <int WS, int BS>
FR_HD void add_shifted10(uint32_t* T, const uint32_t* src) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 18 - WS; i++) {
        uint32_t piece = 0;
        if (i < 10) piece |= src[i] << BS;
        if (BS != 0 && i >= 1 && i - 1 < 10) piece |= src[i - 1] >> (32 - BS);
        c += (uint64_t)T[WS + i] + piece;
        T[WS + i] = (uint32_t)c;
        c >>= 32;
    }
}

// P0, P1, P2 -> canonical Montgomery-form field element  (sum_k a_k*b_k) * 2^-256 mod r
FR_HD Fr kara_finalize(const KAcc& p0, const KAcc& p1, const KAcc& p2) {
    uint32_t a0[10], a1[10], a2[10], mid[10];
    kacc_collapse(p0, a0);
    kacc_collapse(p1, a1);
    kacc_collapse(p2, a2);
    // mid = P1 - P0 - P2 >= 0 (it is the sum of the cross terms lo*hi' + hi*lo')
    int64_t bw = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const int64_t d = (int64_t)a1[i] - (int64_t)a0[i] - (int64_t)a2[i] + bw;
        mid[i] = (uint32_t)d;
        bw = d >> 32;  // arithmetic shift: -2, -1 or 0
    }
    uint32_t T[18];
#pragma unroll
    for (int i = 0; i < 18; i++) T[i] = i < 10 ? a0[i] : 0u;
    add_shifted10<3, 31>(T, mid);  // * 2^127
    add_shifted10<7, 30>(T, a2);   // * 2^254
    return reduce_wide_acc(T);
}

}  // namespace fr
