// Arithmetic of the small-operand tensor-core mat-mul engine (matmul_tc.cu, TcSmall), host-callable so that the digit
// extraction, the signed carry and the Montgomery encode are unit-tested on the CPU build box (tests/test_fr_host.py).
//
// Quantized fixed-point inputs (ZkMatrix::new, reference src/matrix/mod.rs:230-252: round(x * 2^P), negatives as
// r - |q|) are SMALL signed integers in standard form: |q| < 2^(P+7) for |x| < 128.  For such operands
//     sum_k a_ik * b_kj  (an integer below 2^(2*70+32) in magnitude)
// is computed exactly from 9 x 9 signed byte digits per product (81 s8 x s8 multiply-adds) instead of the 32 x 32 byte
// planes (1024 u8 x u8) of the full-width Montgomery representation, and Montgomery-encoded ONCE per C element.  The
// result is the same canonical field element -- the same 32 bytes -- as the full-width engine and the reference's
// `elem += a*b` chain (src/matrix/mod.rs:525-535) produce.
//
// Balanced digits: s = sum_{p<9} d_p * 256^p with d_p in [-128, 127] exists for every |s| <= 2^70 and is read off the
// bytes of s + C, C = sum_p 128 * 256^p:  d_p = byte_p(s + C) - 128, i.e. the s8 bit pattern is byte_p(s + C) XOR 0x80.
#pragma once
#include "fr_fast.cuh"

namespace fr {

constexpr int SMALL_DIGITS = 9;                 // signed byte digits per operand
constexpr int SMALL_BITS = 70;                  // |operand| < 2^70
constexpr uint32_t SMALL_C0 = 0x80808080u, SMALL_C1 = 0x80808080u, SMALL_C2 = 0x00000080u;  // C = 0x80 repeated 9 times

// am: canonical Montgomery form.  Returns true and t = (s + C) mod 2^96 (s = the signed standard-form value of am)
// when |s| < 2^70; returns false (t unspecified) otherwise.
FR_HD bool small_biased(const Fr& am, uint32_t* t) {
    const Fr x = mont_reduce_fast(am);          // canonical integer in [0, r)
    uint32_t m[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = modulus(i);
    sub_n<8>(y, m, x.l);                         // r - x  (> 0)
    const uint32_t hi_x = x.l[3] | x.l[4] | x.l[5] | x.l[6] | x.l[7] | (x.l[2] >> (SMALL_BITS - 64));
    const uint32_t hi_y = y[3] | y[4] | y[5] | y[6] | y[7] | (y[2] >> (SMALL_BITS - 64));
    const bool pos = hi_x == 0, neg = hi_y == 0;
    // s mod 2^96: x itself, or -(r - x)
    uint32_t s0 = x.l[0], s1 = x.l[1], s2 = x.l[2];
    if (!pos) {
        uint64_t c = (uint64_t)(~y[0]) + 1u;
        s0 = (uint32_t)c;
        c = (uint64_t)(~y[1]) + (c >> 32);
        s1 = (uint32_t)c;
        c = (uint64_t)(~y[2]) + (c >> 32);
        s2 = (uint32_t)c;
    }
    uint64_t c = (uint64_t)s0 + SMALL_C0;
    t[0] = (uint32_t)c;
    c = (uint64_t)s1 + SMALL_C1 + (c >> 32);
    t[1] = (uint32_t)c;
    c = (uint64_t)s2 + SMALL_C2 + (c >> 32);
    t[2] = (uint32_t)c;
    return pos || neg;
}
// The same test and value with a third of the multiplications (32 instead of 80 wide products per element; the operand
// split kernels are bound by exactly this).  For a small positive x, am = x * 2^256 mod r means x * 2^256 = am + j * r
// for the integer j = (x * 2^256 - am) / r < x * 2^256 / r < 2^73, and j is determined by j = -am * r^-1 (mod 2^96): three
// Montgomery steps on the low three limbs only (8 products).  For a small negative x the same holds for r - am with
// j' = -(r - am) * r^-1 = -1 - j (mod 2^96), the bitwise complement -- so the sign hypothesis costs nothing.  One exact
// 3 x 8-limb product then gives S = w + j * r, and w is the Montgomery form of a small |x| IF AND ONLY IF the low 256 bits
// of S vanish and S >> 256 < 2^70: no false positives (the identity proves it), no false negatives (the bound on j).
FR_HD bool small_biased_fast(const Fr& am, uint32_t* t) {
    const uint32_t r0 = modulus(0), r1 = modulus(1), r2 = modulus(2);
    // j+ = m0 + m1 * 2^32 + m2 * 2^64: Montgomery multipliers of the low 96 bits
    const uint32_t m0 = am.l[0] * INV32;
    uint64_t c = (uint64_t)m0 * r0 + am.l[0];              // low word becomes 0
    c = (c >> 32) + (uint64_t)m0 * r1 + am.l[1];
    uint32_t t1 = (uint32_t)c;
    c = (c >> 32) + (uint64_t)m0 * r2 + am.l[2];
    uint32_t t2 = (uint32_t)c;
    const uint32_t m1 = t1 * INV32;
    c = (uint64_t)m1 * r0 + t1;
    c = (c >> 32) + (uint64_t)m1 * r1 + t2;
    t2 = (uint32_t)c;
    const uint32_t m2 = t2 * INV32;
    const bool pos = m2 < (1u << 9), neg = ~m2 < (1u << 9);
    if (!pos && !neg) return false;
    uint32_t w[8], j[3], m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = modulus(i);
    if (pos) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = am.l[i];
        j[0] = m0; j[1] = m1; j[2] = m2;
    } else {
        sub_n<8>(w, m, am.l);                               // r - am: Montgomery form of |x|
        j[0] = ~m0; j[1] = ~m1; j[2] = ~m2;
    }
    // S = w + j * r  (11 limbs); its low 8 limbs must vanish
    uint32_t S[11];
#pragma unroll
    for (int i = 0; i < 11; i++) S[i] = i < 8 ? w[i] : 0u;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        uint64_t cy = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) {
            cy += (uint64_t)j[i] * m[l] + S[i + l];
            S[i + l] = (uint32_t)cy;
            cy >>= 32;
        }
#pragma unroll
        for (int l = i + 8; l < 11; l++) {
            cy += S[l];
            S[l] = (uint32_t)cy;
            cy >>= 32;
        }
    }
    uint32_t low = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) low |= S[i];
    if (low != 0 || (S[10] >> (SMALL_BITS - 64)) != 0) return false;
    uint32_t s0 = S[8], s1 = S[9], s2 = S[10];             // |x|
    if (!pos) {                                             // -|x| mod 2^96
        uint64_t n = (uint64_t)(~s0) + 1u;
        s0 = (uint32_t)n;
        n = (uint64_t)(~s1) + (n >> 32);
        s1 = (uint32_t)n;
        n = (uint64_t)(~s2) + (n >> 32);
        s2 = (uint32_t)n;
    }
    c = (uint64_t)s0 + SMALL_C0;
    t[0] = (uint32_t)c;
    c = (uint64_t)s1 + SMALL_C1 + (c >> 32);
    t[1] = (uint32_t)c;
    c = (uint64_t)s2 + SMALL_C2 + (c >> 32);
    t[2] = (uint32_t)c;
    return true;
}
// s8 bit pattern of balanced digit p (0 <= p < 9) of a value biased by small_biased
FR_HD uint32_t small_digit_bits(const uint32_t* t, int p) { return ((t[p >> 2] >> ((p & 3) * 8)) & 0xffu) ^ 0x80u; }

// dg[0..ND): signed 32-bit diagonal sums D_d = sum_{p+q=d} sum_k a_p * b_q  ->  T[0..6) = sum_d D_d * 256^d as a 192-bit
// two's-complement integer (ND <= 20: |value| < 2^(8*19+32)).
template <int ND>
FR_HD void carry_signed(const uint32_t* dg, uint32_t* T) {
    static_assert(ND >= 1 && ND <= 20, "six words hold at most 20 diagonals");
    long long cy = 0;
#pragma unroll
    for (int w = 0; w < 6; w++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int d = 4 * w + r;
            if (d < ND) cy += (long long)(int32_t)dg[d] * (long long)(1 << (8 * r));  // |term| < 2^55, |cy| < 2^58
        }
        T[w] = (uint32_t)cy;
        cy >>= 32;  // arithmetic shift: floor division, the sign travels with the carry
    }
}

// T: 192-bit two's-complement integer, |T| < r  ->  Montgomery form of (T mod r), canonical
FR_HD Fr signed6_to_mont(const uint32_t* T) {
    const uint32_t neg = T[5] >> 31, mask = 0u - neg;
    Fr mag = zero();
    uint64_t c = neg;
#pragma unroll
    for (int w = 0; w < 6; w++) {
        c += (uint64_t)(T[w] ^ mask);
        mag.l[w] = (uint32_t)c;
        c >>= 32;
    }
    const Fr res = to_mont_fast(mag);
    return neg ? sub_fast(zero(), res) : res;  // (0 - res) mod r; sub_fast keeps 0 at 0
}

}  // namespace fr
