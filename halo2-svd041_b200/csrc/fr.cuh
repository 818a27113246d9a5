// BN254 scalar field (halo2curves bn256::Fr) arithmetic for sm_100a.
//
// Wire format everywhere: 4 x u64 little-endian limbs of x * 2^256 mod r, canonical (< r) --
// reinterpreted on the GPU as 8 x u32 little-endian limbs, no conversion.
//
// This header is written twice over on purpose:
//   * everything here is portable `__host__ __device__` C++ (64-bit intermediates), so the exact
//     code the slow/epilogue paths run on the GPU is unit-tested on the CPU build box;
//   * the mat-mul inner loop uses the PTX carry-chain primitives in fr_acc.cuh, whose host
//     fallback is a bit-identical C restatement of the same chain (tested the same way).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#else
#define FR_HD inline
#endif

namespace fr {

struct alignas(16) Fr {
    uint32_t l[8];
};

// ---- constants (derived numerically; checked against halo2curves' published values in tests) ----
FR_HD constexpr uint32_t modulus(int i) {
    constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
}
// R = 2^256 mod r  (Montgomery form of 1)
FR_HD constexpr uint32_t mont_one(int i) {
    constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return m[i];
}
// R^2 mod r
FR_HD constexpr uint32_t mont_r2(int i) {
    constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
}
// 2^480 mod r  (folds accumulator limbs >= 15 before the final Montgomery reduction)
FR_HD constexpr uint32_t pow2_480(int i) {
    constexpr uint32_t m[8] = {0x8f8fc595u, 0x41a9cbe4u, 0x0a58dc0fu, 0x9f36d601u,
                               0x440ad260u, 0xa6132126u, 0x093f81ddu, 0x2ab23b99u};
    return m[i];
}
constexpr uint32_t INV32 = 0xefffffffu;  // -r^-1 mod 2^32

// ---- multi-limb helpers -------------------------------------------------------------------------
template <int N>
FR_HD uint32_t add_n(uint32_t* o, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        c += (uint64_t)a[i] + b[i];
        o[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
template <int N>
FR_HD uint32_t sub_n(uint32_t* o, const uint32_t* a, const uint32_t* b) {
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - borrow;
        o[i] = (uint32_t)d;
        borrow = (uint32_t)(d >> 32) & 1u;
    }
    return borrow;
}
// o = a - r if a >= r else a   (a < 2r)
FR_HD void cond_sub_mod(uint32_t* a) {
    uint32_t t[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = modulus(i);
    uint32_t borrow = sub_n<8>(t, a, m);
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}
FR_HD bool is_canonical(const Fr& a) {
    uint32_t t[8], m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = modulus(i);
    return sub_n<8>(t, a.l, m) != 0;
}
FR_HD bool is_zero(const Fr& a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.l[i];
    return x == 0;
}
FR_HD Fr zero() {
    Fr z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.l[i] = 0;
    return z;
}
FR_HD Fr one() {
    Fr z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.l[i] = mont_one(i);
    return z;
}

// ---- field add / sub (same in the integer and the Montgomery domain) --------------------------------
FR_HD Fr add(const Fr& a, const Fr& b) {
    Fr o;
    add_n<8>(o.l, a.l, b.l);  // < 2r < 2^255: no carry out
    cond_sub_mod(o.l);
    return o;
}
FR_HD Fr sub(const Fr& a, const Fr& b) {
    Fr o;
    uint32_t borrow = sub_n<8>(o.l, a.l, b.l);
    uint32_t m[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = modulus(i);
    add_n<8>(t, o.l, m);
#pragma unroll
    for (int i = 0; i < 8; i++) o.l[i] = borrow ? t[i] : o.l[i];
    return o;
}

// ---- Montgomery multiplication ------------------------------------------------------------------------
// t[0..NA+8) = a[0..NA) * b[0..8)   (schoolbook, 64-bit mac: never overflows)
template <int NA>
FR_HD void mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < NA + 8; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a[i] * b[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
}
// Montgomery reduction: t (16 limbs, value < r * 2^256) -> t * 2^-256 mod r, canonical.
FR_HD Fr redc(uint32_t* t) {
    uint32_t top = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * INV32;
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)m * modulus(j) + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        c += (uint64_t)t[i + 8] + top;
        t[i + 8] = (uint32_t)c;
        top = (uint32_t)(c >> 32);
    }
    // value < 2r < 2^255  =>  top == 0
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.l[i] = t[8 + i];
    cond_sub_mod(o.l);
    return o;
}
FR_HD Fr mont_mul(const Fr& a, const Fr& b) {
    uint32_t t[16];
    mul_wide<8>(t, a.l, b.l);
    return redc(t);
}
// integer (canonical, < r) -> Montgomery form
FR_HD Fr to_mont(const Fr& x) {
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = mont_r2(i);
    return mont_mul(x, r2);
}
// Montgomery form -> canonical integer
FR_HD Fr from_mont(const Fr& a) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t[i] = a.l[i];
        t[8 + i] = 0;
    }
    return redc(t);
}

// ---- lazy-accumulator finalisation -------------------------------------------------------------------
// T = sum over k of a_k * b_k as an 18-limb integer (T < 2^540, i.e. k < 2^32 products of canonical
// operands).  Returns T * 2^-256 mod r, canonical: exactly what the reference's chain of
// `elem += a*b` (reference src/matrix/mod.rs:530) leaves in `elem`, since the canonical representative
// is unique (SURVEY.md A.1 "bit-exactness consequence").
FR_HD Fr reduce_wide_acc(const uint32_t* T) {
    // fold limbs 15..16 (T >> 480 < 2^60; limb 17 is zero for k < 2^32):  T' = T_lo + T_hi * (2^480 mod r)
    uint32_t hi[2] = {T[15], T[16]};
    uint32_t c480[8], prod[10], t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) c480[i] = pow2_480(i);
    mul_wide<2>(prod, hi, c480);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        c += (uint64_t)(i < 15 ? T[i] : 0u) + (i < 10 ? prod[i] : 0u);
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    // T' < 2^480 + 2^314 < r * 2^256  =>  redc precondition holds, c == 0
    return redc(t);
}

// ---- integer helpers on canonical 8-limb values (rescale path) -----------------------------------------
// (x >> s) for 0 <= s < 256.  Staged conditional word moves (s is warp-uniform in practice) so that
// no limb array is ever indexed dynamically (that would force it into local memory).
FR_HD Fr shr(const Fr& x, int s) {
    uint32_t t[9];
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = x.l[i];
    t[8] = 0;
    if (s & 128) {
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = (i + 4 < 8) ? t[i + 4] : 0u;
    }
    if (s & 64) {
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = (i + 2 < 8) ? t[i + 2] : 0u;
    }
    if (s & 32) {
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = (i + 1 < 8) ? t[i + 1] : 0u;
    }
    const int b = s & 31;
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) o.l[i] = (uint32_t)((((uint64_t)t[i + 1] << 32) | t[i]) >> b);
    return o;
}
// x mod 2^bits for 0 <= bits <= 256
FR_HD Fr low_bits(const Fr& x, int bits) {
    Fr o;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int rem = bits - 32 * i;
        o.l[i] = rem >= 32 ? x.l[i] : (rem <= 0 ? 0u : (x.l[i] & ((1u << rem) - 1u)));
    }
    return o;
}
FR_HD Fr pow2(int s) {
    Fr o = zero();
#pragma unroll
    for (int i = 0; i < 8; i++) o.l[i] = (i == (s >> 5)) ? (1u << (s & 31)) : 0u;
    return o;
}

}  // namespace fr
