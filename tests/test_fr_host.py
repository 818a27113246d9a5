"""CPU tests of the PRODUCT's field-arithmetic headers (csrc/fr.cuh, csrc/fr_acc.cuh) compiled for
the host: the portable code is what the GPU runs outside the PTX carry chains, and the host
fallback of the chain mirrors the PTX bit for bit (positions, carry counters, final fold + REDC)."""
import ctypes as ct
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import ROOT, raw_limbs

R = po.R_MOD


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("frhost") / "libfrhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "fr_host_shim.cpp")])
    return ct.CDLL(so)


def P(a):
    return a.ctypes.data_as(ct.c_void_p)


def unraw(a):
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a.reshape(-1, 4)]


def _vals():
    rng = random.Random(7)
    edge = [0, 1, 2, R - 1, R - 2, 1 << 253, (1 << 252) - 1, po.MONT_R, po.MONT_R2, (R - 1) // 2]
    return edge + [rng.randrange(R) for _ in range(200)]


def test_field_ops_match_bigint(shim):
    vals = _vals()
    n = len(vals)
    a, b = raw_limbs(vals), raw_limbs(vals[::-1])
    o = np.zeros_like(a)
    shim.hs_mont_mul(P(a), P(b), P(o), n)
    assert unraw(o) == [x * y * po.MONT_RINV % R for x, y in zip(vals, vals[::-1])]
    shim.hs_add(P(a), P(b), P(o), n)
    assert unraw(o) == [(x + y) % R for x, y in zip(vals, vals[::-1])]
    shim.hs_sub(P(a), P(b), P(o), n)
    assert unraw(o) == [(x - y) % R for x, y in zip(vals, vals[::-1])]
    shim.hs_to_mont(P(a), P(o), n)
    assert unraw(o) == [po.to_mont(x) for x in vals]
    shim.hs_from_mont(P(a), P(o), n)
    assert unraw(o) == [po.from_mont(x) for x in vals]
    one = np.zeros((1, 4), dtype=np.uint64)
    shim.hs_one(P(one))
    assert unraw(one) == [po.MONT_R]
    assert shim.hs_is_canonical(P(raw_limbs([R - 1]))) == 1
    assert shim.hs_is_canonical(P(raw_limbs([R]))) == 0
    assert shim.hs_is_canonical(P(raw_limbs([(1 << 256) - 1]))) == 0


def test_carry_chain_primitives_match_bigint(shim):
    """fr_fast.cuh: the two-limbs-per-step CIOS in even/odd 64-bit columns, the single-limb Montgomery
    step and the modular add/sub, on edge values and 20k random pairs."""
    rng = random.Random(11)
    edge = [0, 1, 2, R - 1, R - 2, 1 << 253, (1 << 252) - 1, po.MONT_R, po.MONT_R2, (R - 1) // 2,
            (1 << 32) - 1, (1 << 64) - 1, ((1 << 254) - 1) % R, R - (1 << 32), 0x30644e72 << 224]
    xs = [x for x in edge for _ in edge] + [rng.randrange(R) for _ in range(20000)]
    ys = [y for _ in edge for y in edge] + [rng.randrange(R) for _ in range(20000)]
    n = len(xs)
    a, b = raw_limbs(xs), raw_limbs(ys)
    o = np.zeros_like(a)
    shim.hs_mont_mul_fast(P(a), P(b), P(o), n)
    assert unraw(o) == [x * y * po.MONT_RINV % R for x, y in zip(xs, ys)]
    shim.hs_add_fast(P(a), P(b), P(o), n)
    assert unraw(o) == [(x + y) % R for x, y in zip(xs, ys)]
    shim.hs_sub_fast(P(a), P(b), P(o), n)
    assert unraw(o) == [(x - y) % R for x, y in zip(xs, ys)]
    shim.hs_to_mont_fast(P(a), P(o), n)
    assert unraw(o) == [po.to_mont(x) for x in xs]
    shim.hs_from_mont_fast(P(a), P(o), n)
    assert unraw(o) == [po.from_mont(x) for x in xs]
    ls = [0, 1, 2, (1 << 19) - 1, (1 << 31), (1 << 32) - 1] + [rng.randrange(1 << 32) for _ in range(5000)]
    cs = [rng.randrange(R) for _ in ls]
    cs[:6] = [R - 1, R - 1, 0, R - 1, R - 1, R - 1]
    larr = np.array(ls, dtype=np.uint32)
    c = raw_limbs(cs)
    o = np.zeros_like(c)
    shim.hs_mont_mul_small(P(larr), P(c), P(o), len(ls))
    inv32 = pow(1 << 32, -1, R)
    assert unraw(o) == [l * x * inv32 % R for l, x in zip(ls, cs)]


def _small_sums(shim, ls, rhos):
    n = len(ls)
    larr = np.array(ls, dtype=np.uint32)
    rho = raw_limbs(rhos)
    phi = np.array([(x << 64) // R for x in rhos], dtype=np.uint64)
    o = np.zeros_like(rho)
    shim.hs_small_sums(P(larr), P(rho), P(phi), P(o), n)
    return unraw(o)


@pytest.mark.parametrize("lb", [1, 8, 12, 19, 27])
def test_small_sum_running_sums_are_canonical_partial_sums(shim, lb):
    """fr::SmallSum (the range-check running sums of the rescale / range-check kernels): every partial sum of
    l_j * M(2^(lb*j)) equals the exact residue, for random limbs, all-ones limbs (the largest sums) and zeros."""
    rng = random.Random(lb)
    npos = min(32, 253 // lb, (1 << 32) >> lb)
    rhos = [(1 << (lb * i)) * (1 << 256) % R for i in range(npos)]
    cases = [[rng.randrange(1 << lb) for _ in range(npos)] for _ in range(200)]
    cases += [[(1 << lb) - 1] * npos, [0] * npos, [1] * npos]
    for ls in cases:
        want, acc = [], 0
        for l, rho in zip(ls, rhos):
            acc = (acc + l * rho) % R
            want.append(acc)
        assert _small_sums(shim, ls, rhos) == want


def test_small_sum_quotient_estimate_edge(shim):
    """The 64-bit fixed-point quotient estimate can be one short only when the fraction of F is within 2^-32 of 1; the code
    takes the exact compare-and-subtract whenever the top 20 fraction bits are ones.  Search limb pairs (l0, l1) that
    land there (about 2^-20 of all pairs) and check them, plus their neighbours, against exact arithmetic."""
    lb = 19
    rhos = [(1 << (lb * i)) * (1 << 256) % R for i in range(2)]
    phi = [(x << 64) // R for x in rhos]
    l0 = np.arange(1 << lb, dtype=np.uint64)
    hits = []
    for l1 in range(1, 97):
        f = l0 * np.uint64(phi[0]) + np.uint64((l1 * phi[1]) % (1 << 64))     # mod 2^64: the fraction of F
        for a in np.nonzero((f >> np.uint64(44)) == np.uint64(0xFFFFF))[0]:
            hits.append((int(a), l1))
    assert len(hits) >= 10
    for a, l1 in hits:
        for ls in ([a, l1], [max(a - 1, 0), l1], [min(a + 1, (1 << lb) - 1), l1]):
            want = [ls[0] * rhos[0] % R, (ls[0] * rhos[0] + ls[1] * rhos[1]) % R]
            assert _small_sums(shim, ls, rhos) == want


def test_shift_and_mask_helpers(shim):
    vals = _vals()
    n = len(vals)
    a = raw_limbs(vals)
    o = np.zeros_like(a)
    for s in [0, 1, 19, 31, 32, 33, 63, 64, 65, 127, 128, 129, 190, 255]:
        shim.hs_shr(P(a), s, P(o), n)
        assert unraw(o) == [x >> s for x in vals]
        shim.hs_low_bits(P(a), s, P(o), n)
        assert unraw(o) == [x & ((1 << s) - 1) for x in vals]
        shim.hs_pow2(s, P(o))
        assert unraw(o[:1]) == [1 << s]
    shim.hs_low_bits(P(a), 256, P(o), n)
    assert unraw(o) == vals


@pytest.mark.parametrize("k", [1, 2, 3, 17, 64, 1000])
def test_lazy_dot_product(shim, k):
    rng = random.Random(k)
    A = [rng.randrange(R) for _ in range(k)]
    B = [rng.randrange(R) for _ in range(k)]
    o = np.zeros((1, 4), dtype=np.uint64)
    shim.hs_lazy_dot(P(raw_limbs(A)), P(raw_limbs(B)), k, P(o))
    assert unraw(o) == [sum(x * y for x, y in zip(A, B)) * po.MONT_RINV % R]


def test_lazy_accumulator_headroom(shim):
    """Largest operands repeated k times: exercises every carry counter and the 2^480 fold."""
    mx = raw_limbs([R - 1])
    o = np.zeros((1, 4), dtype=np.uint64)
    for k in [1, 7, 100000, 3000000]:
        shim.hs_lazy_repeat(P(mx), P(mx), k, P(o))
        assert unraw(o) == [k * (R - 1) * (R - 1) * po.MONT_RINV % R]
    top = raw_limbs([(1 << 256) - 1])  # non-canonical limbs: still exact as integers
    shim.hs_lazy_repeat(P(top), P(top), 1000, P(o))
    assert unraw(o) == [1000 * ((1 << 256) - 1) ** 2 * po.MONT_RINV % R]


def test_lazy_matmul_ragged(shim):
    rng = random.Random(5)
    n, k, m = 5, 9, 3
    A = [rng.randrange(R) for _ in range(n * k)]
    B = [rng.randrange(R) for _ in range(k * m)]
    C = np.zeros((n * m, 4), dtype=np.uint64)
    shim.hs_lazy_matmul(P(raw_limbs(A)), P(raw_limbs(B)), P(C), n, k, m)
    exp = [sum(A[i * k + t] * B[t * m + j] for t in range(k)) * po.MONT_RINV % R
           for i in range(n) for j in range(m)]
    assert unraw(C) == exp


@pytest.mark.parametrize("k", [1, 2, 3, 17, 64, 1000])
def test_karatsuba_lazy_dot_product(shim, k):
    """fr_kara.cuh: split at bit 127, three lazy 4x4-limb product sums, recombination + reduction in the epilogue."""
    rng = random.Random(100 + k)
    edge = [0, 1, R - 1, (1 << 127) - 1, 1 << 127, (1 << 253), ((1 << 254) - 1) % R, (1 << 128) - 1]
    xs = ([rng.choice(edge) for _ in range(k // 2)] + [rng.randrange(R) for _ in range(k)])[:k]
    ys = ([rng.choice(edge) for _ in range(k // 3)] + [rng.randrange(R) for _ in range(k)])[:k]
    a, b = raw_limbs(xs), raw_limbs(ys)
    o = np.zeros((1, 4), dtype=np.uint64)
    shim.hs_kara_dot(P(a), P(b), k, P(o))
    assert unraw(o) == [sum(x * y for x, y in zip(xs, ys)) * po.MONT_RINV % R]


def test_karatsuba_accumulator_headroom(shim):
    """worst-case operands (r - 1) repeated: the 4x4 accumulators and their carry counters must not wrap"""
    a = raw_limbs([R - 1])
    o = np.zeros((1, 4), dtype=np.uint64)
    for k in (1, 4096, 100000):
        shim.hs_kara_repeat(P(a), P(a), k, P(o))
        assert unraw(o) == [k * (R - 1) * (R - 1) * po.MONT_RINV % R]


def test_mont_reduce_fast_matches_bigint(shim):
    vals = _vals()
    a = raw_limbs(vals)
    o = np.zeros_like(a)
    shim.hs_mont_reduce_fast(P(a), P(o), len(vals))
    assert unraw(o) == [po.from_mont(x) for x in vals]


def _mont_signed(vals):
    return raw_limbs([po.to_mont(v % R) for v in vals])


@pytest.mark.parametrize("k", [1, 7, 64, 1024])
def test_small_operand_engine_arithmetic(shim, k):
    """tc_small.cuh: balanced signed byte digits -> 17 diagonal sums -> signed carry -> one Montgomery encode gives the
    same canonical element as sum a*b over the field, incl. the extremes +-(2^70 - 1), 0, +-1 and mixed signs."""
    rng = random.Random(k)
    lim = (1 << 70) - 1
    edge = [0, 1, -1, lim, -lim, 127, 128, -128, -129, 255, 256, -255, -256, (1 << 63), -(1 << 63), (1 << 64) - 1]
    av = [rng.choice(edge) if rng.random() < 0.3 else rng.randrange(-lim, lim + 1) for _ in range(k)]
    bv = [rng.choice(edge) if rng.random() < 0.3 else rng.randrange(-lim, lim + 1) for _ in range(k)]
    o = np.zeros((1, 4), dtype=np.uint64)
    assert shim.hs_small_dot(P(_mont_signed(av)), P(_mont_signed(bv)), k, P(o)) == 1
    assert unraw(o) == [po.to_mont(sum(x * y for x, y in zip(av, bv)) % R)]
    # worst case for the accumulators: every product +-(2^70-1)^2
    av, bv = [lim] * k, [-lim if i % 2 else lim for i in range(k)]
    assert shim.hs_small_dot(P(_mont_signed(av)), P(_mont_signed(bv)), k, P(o)) == 1
    assert unraw(o) == [po.to_mont(sum(x * y for x, y in zip(av, bv)) % R)]


def test_small_operand_range_detection(shim):
    lim = 1 << 70
    for v, ok in [(0, 1), (lim - 1, 1), (-(lim - 1), 1), (lim, 0), (-lim, 0), (lim + 5, 0), (R // 2, 0), (-(R // 2), 0),
                  ((1 << 200) + 3, 0), (1 << 96, 0), (-(1 << 96), 0)]:
        assert shim.hs_small_ok(P(_mont_signed([v]))) == ok, v


def test_small_operand_fast_detection_agrees_with_full_reduction(shim):
    """tc_small.cuh small_biased_fast (32 products: three Montgomery steps + one 3 x 8-limb product) == small_biased (full
    reduction) on small values of both signs, the range boundaries, near-misses and arbitrary field elements."""
    rng = random.Random(99)
    lim = 1 << 70
    small = [0, 1, -1, lim - 1, -(lim - 1), 2, -2, (1 << 64), -(1 << 64), (1 << 69) + 12345, 255, -256]
    small += [rng.randrange(-lim + 1, lim) for _ in range(3000)]
    near = [lim, -lim, lim + 1, -(lim + 1), (1 << 71), -(1 << 71), (1 << 72) + 5, -(1 << 73), (1 << 96), (1 << 95) - 1, R // 2]
    big = [rng.randrange(R) for _ in range(3000)]
    # Montgomery forms of values whose low multipliers happen to look small-ish: x * R with x just above the range
    near += [rng.randrange(lim, lim << 6) * rng.choice((1, -1)) for _ in range(500)]
    arr = _mont_signed(small)
    assert shim.hs_small_fast_vs_full(P(arr), len(small)) == len(small)
    arr = _mont_signed(near)
    assert shim.hs_small_fast_vs_full(P(arr), len(near)) == 0
    arr = raw_limbs(big)                       # raw limb patterns: arbitrary Montgomery-form elements
    assert shim.hs_small_fast_vs_full(P(arr), len(big)) == 0
    mixed = _mont_signed(small[:50] + near[:50])
    assert shim.hs_small_fast_vs_full(P(mixed), 100) == 50


def test_mont_mul_fast_x2_matches_bigint(shim):
    vals = _vals()
    vals = vals[: len(vals) // 2 * 2]
    a, b = raw_limbs(vals), raw_limbs(vals[::-1])
    o = np.zeros_like(a)
    shim.hs_mont_mul_fast_x2(P(a), P(b), P(o), len(vals))
    assert unraw(o) == [x * y * po.MONT_RINV % R for x, y in zip(vals, vals[::-1])]
