// Lazy (unreduced) accumulation of 8x8-limb products for the Fr mat-mul inner loop.
//
// One accumulator holds sum_k a_k * b_k as a redundant 18-limb integer:
//   e[8]  : 64-bit column sums at even limb positions  (pair p = limbs 2p, 2p+1)
//   o[7]  : 64-bit column sums at odd  limb positions  (pair p = limbs 2p+1, 2p+2)
//   ce[5] : carry counters out of the even chains, weight 2^(32*(8+2q))
//   co[4] : carry counters out of the odd  chains, weight 2^(32*(9+2q))
// A product a_i*b_j lands in the even set when i+j is even, else in the odd set, so every
// mad.lo.cc / madc.hi.cc pair targets an aligned register pair and ptxas fuses it into ONE
// `IMAD.WIDE.U32[.X]` with the carry travelling in a predicate: 64 IMAD.WIDE + 16 IADD3.X per
// Fr multiply-add, no Montgomery reduction inside the k loop (done once in fr::reduce_wide_acc).
#pragma once
#include "fr.cuh"

namespace fr {

struct WideAcc {
    uint64_t e[8];   // pair p = limbs 2p, 2p+1
    uint64_t o[7];   // pair p = limbs 2p+1, 2p+2
    uint32_t ce[5];
    uint32_t co[4];
};

FR_HD void acc_clear(WideAcc& w) {
#pragma unroll
    for (int i = 0; i < 8; i++) w.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) w.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) w.ce[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) w.co[i] = 0;
}

// d[0..4) (four 64-bit columns) += {a0,a1,a2,a3} * b with one carry chain; cnt += carry out.
// The columns are kept as 64-bit values and split/joined with mov.b64 inside the asm block so that
// ptxas sees aligned register PAIRS for the whole k loop (otherwise it inserts IMAD.MOVs to re-pair
// 32-bit halves, which cost slots on the very pipe the kernel is bound by).
FR_HD void chain4(uint64_t* d, uint32_t& cnt, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                  uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
        "mov.b64 {l0, h0}, %0;\n\t"
        "mov.b64 {l1, h1}, %1;\n\t"
        "mov.b64 {l2, h2}, %2;\n\t"
        "mov.b64 {l3, h3}, %3;\n\t"
        "mad.lo.cc.u32   l0, %5, %9, l0;\n\t"
        "madc.hi.cc.u32  h0, %5, %9, h0;\n\t"
        "madc.lo.cc.u32  l1, %6, %9, l1;\n\t"
        "madc.hi.cc.u32  h1, %6, %9, h1;\n\t"
        "madc.lo.cc.u32  l2, %7, %9, l2;\n\t"
        "madc.hi.cc.u32  h2, %7, %9, h2;\n\t"
        "madc.lo.cc.u32  l3, %8, %9, l3;\n\t"
        "madc.hi.cc.u32  h3, %8, %9, h3;\n\t"
        "addc.u32        %4, %4, 0;\n\t"
        "mov.b64 %0, {l0, h0};\n\t"
        "mov.b64 %1, {l1, h1};\n\t"
        "mov.b64 %2, {l2, h2};\n\t"
        "mov.b64 %3, {l3, h3};\n\t"
        "}"
        : "+l"(d[0]), "+l"(d[1]), "+l"(d[2]), "+l"(d[3]), "+r"(cnt)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    const uint32_t a[4] = {a0, a1, a2, a3};
    uint32_t carry = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a[i] * b) + d[i] + carry;
        d[i] = (uint64_t)t;
        carry = (uint32_t)(t >> 64);
    }
    cnt += carry;
#endif
}

// w += a * b   (a, b: 8 x u32 limbs each)
FR_HD void mul_acc(WideAcc& w, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i0 = j & 1;         // a-limbs with i+j even
        const int pe = (i0 + j) >> 1; // first even pair touched, 0..4
        chain4(&w.e[pe], w.ce[pe], a[i0], a[i0 + 2], a[i0 + 4], a[i0 + 6], b[j]);
        const int i1 = 1 - i0;        // a-limbs with i+j odd
        const int po = j >> 1;        // first odd pair touched, 0..3
        chain4(&w.o[po], w.co[po], a[i1], a[i1 + 2], a[i1 + 4], a[i1 + 6], b[j]);
    }
}

// Collapse the redundant form into a plain 18-limb integer T (T < 2^540 for < 2^31 products).
FR_HD void acc_collapse(const WideAcc& w, uint32_t* T) {
    uint64_t col[18];
#pragma unroll
    for (int i = 0; i < 18; i++) col[i] = 0;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        col[2 * p] += (uint32_t)w.e[p];
        col[2 * p + 1] += (uint32_t)(w.e[p] >> 32);
    }
#pragma unroll
    for (int p = 0; p < 7; p++) {
        col[2 * p + 1] += (uint32_t)w.o[p];
        col[2 * p + 2] += (uint32_t)(w.o[p] >> 32);
    }
#pragma unroll
    for (int q = 0; q < 5; q++) col[8 + 2 * q] += w.ce[q];
#pragma unroll
    for (int q = 0; q < 4; q++) col[9 + 2 * q] += w.co[q];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) {
        c += col[i];
        T[i] = (uint32_t)c;
        c >>= 32;
    }
}

// sum_k a_k*b_k  ->  canonical Montgomery-form field element (see fr::reduce_wide_acc)
FR_HD Fr acc_finalize(const WideAcc& w) {
    uint32_t T[18];
    acc_collapse(w, T);
    return reduce_wide_acc(T);
}

}  // namespace fr
