"""Python big-int ORACLE for the ZkMatrix / ZkVector witness path.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and only as the checker.

PARITY STATUS: **parity unpinned** for the third-party boundary.  The reference
(/root/reference, Rust) cannot be built here (no cargo/rustc, un-vendored
crates) and ships no golden vectors for this path.  What
IS pinned:
  * BN254 Fr constants (modulus, R, R^2, INV) -- checked in tests against the
    values halo2curves publishes;
  * the in-tree loop semantics of src/matrix/mod.rs (cited per function);
  * the halo2-base 0.4.1 cell layout, validated against the advice / lookup
    cell-count formulas of the reference README.md:67 (135N^2 / 26N^2 at P=32,
    201N^2 / 48N^2 at P=63) in tests/test_oracle.py.
Unpinned: FixedPointChip041::signed_div_scale's shift / a_num_bits constants
(runtime parameters here, defaults S=3P, A=4P) and qsqrt.

All values in this file are *standard-form* Python ints in [0, r).  The wire
format (halo2curves bn256::Fr: 4 x u64 little-endian limbs of x*2^256 mod r) is
produced/consumed only by pack_mont / unpack_mont.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

# --- BN254 scalar field (halo2curves bn256::Fr), SURVEY.md A.1 -------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
MONT_R = (1 << 256) % R_MOD
MONT_R2 = (MONT_R * MONT_R) % R_MOD
MONT_RINV = pow(MONT_R, -1, R_MOD)
MONT_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
MASK64 = (1 << 64) - 1


def to_mont(x: int) -> int:
    return (x % R_MOD) * MONT_R % R_MOD


def from_mont(x: int) -> int:
    return x * MONT_RINV % R_MOD


def pack_mont(values: Sequence[int]) -> np.ndarray:
    """standard-form ints -> uint64[len, 4] Montgomery little-endian limbs."""
    out = np.empty((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = to_mont(v)
        for j in range(4):
            out[i, j] = (m >> (64 * j)) & MASK64
    return out


def unpack_mont(arr: np.ndarray) -> List[int]:
    """uint64[..., 4] Montgomery limbs -> flat list of standard-form ints."""
    a = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in a:
        m = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        if m >= R_MOD:
            raise ValueError("non-canonical Fr limbs")
        out.append(from_mont(m))
    return out


def signed(x: int) -> int:
    """Field element -> signed integer in (-r/2, r/2]."""
    return x if x <= R_MOD // 2 else x - R_MOD


# --- value-level restatements of the reference loops ------------------------

def field_mat_mul(a: Sequence[Sequence[int]], b: Sequence[Sequence[int]]) -> List[List[int]]:
    """reference src/matrix/mod.rs:510-537 (i, j, k order)."""
    assert len(a[0]) == len(b)  # :515
    n, k, m = len(a), len(a[0]), len(b[0])
    c = []
    for i in range(n):
        row = []
        for j in range(m):
            e = 0
            for t in range(k):
                e = (e + a[i][t] * b[t][j]) % R_MOD  # :530
            row.append(e)
        c.append(row)
    return c


def gamma_powers(gamma: int, d: int) -> List[int]:
    """reference src/matrix/mod.rs:316-326: v = (1, g, ..., g^(d-1))."""
    v = [1]
    for _ in range(1, d):
        v.append(v[-1] * gamma % R_MOD)
    return v


def mat_vec_prefix(a: Sequence[Sequence[int]], v: Sequence[int]) -> List[List[int]]:
    """Running sums of gate.inner_product(row, v) for every row
    (reference src/matrix/mod.rs:574-599 + halo2-base inner_product, A.2)."""
    assert len(a[0]) == len(v)  # :580
    out = []
    for row in a:
        s, pre = 0, []
        for x, y in zip(row, v):
            s = (s + x * y) % R_MOD
            pre.append(s)
        out.append(pre)
    return out


def quantize(x: float, precision_bits: int) -> int:
    """FixedPointChip041::quantization (A.5; PDF Eq. 11): sign-magnitude
    round-half-up of |x|*2^P, negatives as r - q."""
    y = abs(x) * float(1 << precision_bits)      # exact: scaling by a power of two
    q = int(math.floor(y))
    if y - math.floor(y) >= 0.5:                  # exact: Rust's f64::round (half away from zero), no y + 0.5 tie-to-even
        q += 1
    return q % R_MOD if x >= 0 else (R_MOD - q) % R_MOD


def dequantize(v: int, precision_bits: int) -> float:
    return signed(v) / float(1 << precision_bits)


@dataclass(frozen=True)
class RescaleParams:
    """signed_div_scale constants (S and A are unpinned)."""
    precision_bits: int
    lookup_bits: int
    shift_bits: int = -1   # S; default 3P
    a_num_bits: int = -1   # A; default 4P

    @property
    def S(self) -> int:
        return 3 * self.precision_bits if self.shift_bits < 0 else self.shift_bits

    @property
    def A(self) -> int:
        return 4 * self.precision_bits if self.a_num_bits < 0 else self.a_num_bits

    @property
    def n_d(self) -> int:
        return -(-(self.A - self.precision_bits + 1) // self.lookup_bits)

    @property
    def n_r(self) -> int:
        return -(-(self.precision_bits + 1) // self.lookup_bits)

    @property
    def W(self) -> int:  # Witness values per element: 4n per check_big_less_than_safe, 2 when its n == 1
        def cbl(n):
            return 4 * n if n >= 2 else 2
        return 4 + cbl(self.n_d) + cbl(self.n_r)

    @property
    def cells(self) -> int:  # advice cells per element
        # each check_big_less_than_safe: range_check (3n-2, or 0 if n==1) + 7 + range_check
        def cbl(n):
            rc = 0 if n == 1 else 3 * n - 2
            return 2 * rc + 7
        return 12 + cbl(self.n_d) + cbl(self.n_r)


def _limbs(x: int, n: int, lb: int) -> List[int]:
    """halo2-base decompose_fe_to_u64_limbs: low n chunks of lb bits."""
    return [(x >> (lb * i)) & ((1 << lb) - 1) for i in range(n)]


def _range_check_witness(x: int, n: int, lb: int) -> List[int]:
    """Witness values of RangeChip::range_check(x, n*lb) (A.4): l0, l1, s1, l2, s2, ..."""
    if n == 1:
        return []
    l = _limbs(x, n, lb)
    out = [l[0], l[1]]
    s = l[0] + (l[1] << lb)
    out.append(s % R_MOD)
    for i in range(2, n):
        out.append(l[i])
        s += l[i] << (lb * i)
        out.append(s % R_MOD)
    return out


def _check_big_less_than_safe_witness(x: int, bound: int, lb: int) -> List[int]:
    """Witness values of RangeChip::check_big_less_than_safe(x, bound) (A.4)."""
    n = -(-bound.bit_length() // lb)
    bits = n * lb
    out = _range_check_witness(x, n, lb)
    chk = (x + (1 << bits) - bound) % R_MOD
    out += [chk, (x + (1 << bits)) % R_MOD]
    out += _range_check_witness(chk, n, lb)
    return out


def signed_div_scale_witness(a: int, p: RescaleParams) -> Tuple[int, int, List[int]]:
    """Per-element witness list of FixedPointChip041::signed_div_scale
    (called from reference src/matrix/mod.rs:369 and :104; model SURVEY A.5).
    Returns (q, rem, [W Witness values in cell order])."""
    P, lb = p.precision_bits, p.lookup_bits
    a_shift = (a + (1 << p.S)) % R_MOD
    div, rem = a_shift >> P, a_shift & ((1 << P) - 1)   # div_mod_floor by 2^P
    w = [a_shift, rem, div]
    w += _check_big_less_than_safe_witness(div, (1 << (p.A - P)) + 1, lb)
    w += _check_big_less_than_safe_witness(rem, 1 << P, lb)
    q = (div - (1 << (p.S - P))) % R_MOD
    w.append(q)
    assert len(w) == p.W
    return q, rem, w


def isqrt_fixed(a: int, precision_bits: int) -> int:
    """qsqrt model (parity unpinned): floor(sqrt(a * 2^P))."""
    return math.isqrt(a << precision_bits)


# --- halo2-base 0.4.1 Context / chips cell-layout model (A.2 - A.5) ---------

@dataclass
class AssignedValue:
    value: int
    index: int
    ctx_id: int = 0   # which virtual column (phase-0 / phase-1 context) the cell lives in


class Context:
    """Virtual advice column of one halo2-base Context (A.2)."""

    def __init__(self, ctx_id: int = 0) -> None:
        self.ctx_id = ctx_id
        self.advice: List[int] = []
        self.kind: List[str] = []        # 'W' | 'E' | 'C' per cell
        self.selector: List[bool] = []
        self.copies: List[Tuple[Tuple[int, int], Tuple[int, int]]] = []  # ((ctx, cell), (ctx, cell))
        self.constants: List[Tuple[int, int]] = []  # (cell, value)
        self.lookups: List[int] = []

    def _push(self, cell) -> None:
        kind, x = cell
        idx = len(self.advice)
        if kind == "E":
            self.advice.append(x.value)
            self.copies.append(((x.ctx_id, x.index), (self.ctx_id, idx)))
        elif kind == "C":
            self.advice.append(x % R_MOD)
            self.constants.append((idx, x % R_MOD))
        else:
            self.advice.append(x % R_MOD)
        self.kind.append(kind)
        self.selector.append(False)

    def assign_region(self, cells, gate_offsets) -> int:
        row = len(self.advice)
        for c in cells:
            self._push(c)
        for off in gate_offsets:
            self.selector[row + off] = True
        return row

    def get(self, i: int) -> AssignedValue:
        if i < 0:
            i += len(self.advice)
        return AssignedValue(self.advice[i], i, self.ctx_id)

    def load_witness(self, v: int) -> AssignedValue:
        self._push(("W", v))
        return self.get(-1)

    def load_constant(self, v: int) -> AssignedValue:
        self._push(("C", v))
        return self.get(-1)

    def constrain_equal(self, a: AssignedValue, b: AssignedValue) -> None:
        self.copies.append(((a.ctx_id, a.index), (b.ctx_id, b.index)))

    def witness_values(self) -> List[int]:
        return [v for v, k in zip(self.advice, self.kind) if k == "W"]


def E(x):  # QuantumCell::Existing
    return ("E", x)


def C(x):  # QuantumCell::Constant
    return ("C", x)


def W(x):  # QuantumCell::Witness
    return ("W", x)


class GateChip:
    def add(self, ctx, a, b):
        ctx.assign_region([a, b, C(1), W(_val(a) + _val(b))], [0])
        return ctx.get(-1)

    def sub(self, ctx, a, b):
        ctx.assign_region([W(_val(a) - _val(b)), b, C(1), a], [0])
        return ctx.get(-4)

    def mul(self, ctx, a, b):
        ctx.assign_region([C(0), a, b, W(_val(a) * _val(b))], [0])
        return ctx.get(-1)

    def inner_product(self, ctx, a, b):
        a, b = list(a), list(b)
        assert len(a) == len(b)
        cells, s = [], 0
        if b and b[0][0] == "C" and b[0][1] % R_MOD == 1:
            s = _val(a[0])
            cells.append(a[0])
            a, b = a[1:], b[1:]
        else:
            cells.append(C(0))
        for x, y in zip(a, b):
            s = (s + _val(x) * _val(y)) % R_MOD
            cells += [x, y, W(s)]
        ctx.assign_region(cells, [3 * i for i in range(len(cells) // 3)])
        return ctx.get(-1)

    def is_zero(self, ctx, a: AssignedValue):
        z = 1 if a.value == 0 else 0
        inv = 1 if a.value == 0 else pow(a.value, -1, R_MOD)
        ctx.assign_region([W(z), E(a), W(inv), C(1), C(0), E(a), W(z), C(0)], [0, 4])
        return ctx.get(-2)

    def is_equal(self, ctx, a, b):
        diff = self.sub(ctx, a, b)
        return self.is_zero(ctx, diff)


def _val(cell) -> int:
    kind, x = cell
    return x.value if kind == "E" else x % R_MOD


class RangeChip:
    def __init__(self, lookup_bits: int) -> None:
        self.lookup_bits = lookup_bits
        self.gate = GateChip()

    def range_check(self, ctx, a: AssignedValue, range_bits: int) -> None:
        lb = self.lookup_bits
        n = -(-range_bits // lb)
        rem_bits = range_bits % lb
        if n == 1:
            ctx.lookups.append(a.index)
            last = a
        else:
            limbs = [W(x) for x in _limbs(a.value, n, lb)]
            bases = [C(1 << (lb * i)) for i in range(n)]
            row = len(ctx.advice)
            acc = self.gate.inner_product(ctx, limbs, bases)
            ctx.constrain_equal(a, acc)
            ctx.lookups.append(row)
            for i in range(n - 1):
                ctx.lookups.append(row + 1 + 3 * i)
            last = ctx.get(row + 1 + 3 * (n - 2))
        if rem_bits == 1:
            ctx.assign_region([C(0), E(last), E(last), E(last)], [0])
        elif rem_bits > 1:
            chk = self.gate.mul(ctx, E(last), C(1 << (lb - rem_bits)))
            ctx.lookups.append(chk.index)

    def check_less_than(self, ctx, a, b, num_bits: int) -> None:
        p2 = 1 << num_bits
        ctx.assign_region(
            [W(p2 + _val(a) - _val(b)), b, C(1), W(p2 + _val(a)), C(-p2), C(1), a], [0, 3])
        self.range_check(ctx, ctx.get(-7), num_bits)

    def check_big_less_than_safe(self, ctx, a: AssignedValue, bound: int) -> None:
        lb = self.lookup_bits
        bits = -(-bound.bit_length() // lb) * lb
        self.range_check(ctx, a, bits)
        self.check_less_than(ctx, E(a), C(bound), bits)

    def div_mod(self, ctx, a, b: int, a_num_bits: int):
        div, rem = divmod(_val(a), b)
        ctx.assign_region([W(rem), C(b), W(div), a], [0])
        rem_c, div_c = ctx.get(-4), ctx.get(-2)
        self.check_big_less_than_safe(ctx, div_c, (1 << a_num_bits) // b + 1)
        self.check_big_less_than_safe(ctx, rem_c, b)
        return div_c, rem_c


class FixedPointChip:
    """FixedPointChip041<F, PRECISION_BITS> model (A.5)."""

    def __init__(self, precision_bits: int, lookup_bits: int, shift_bits: int = -1,
                 a_num_bits: int = -1) -> None:
        self.params = RescaleParams(precision_bits, lookup_bits, shift_bits, a_num_bits)
        self.range = RangeChip(lookup_bits)
        self.gate = self.range.gate

    @property
    def P(self) -> int:
        return self.params.precision_bits

    def quantization(self, x: float) -> int:
        return quantize(x, self.P)

    def dequantization(self, v: int) -> float:
        return dequantize(v, self.P)

    def qsub(self, ctx, a: AssignedValue, b: AssignedValue) -> AssignedValue:
        return self.gate.sub(ctx, E(a), E(b))

    def signed_div_scale(self, ctx, a: AssignedValue):
        p = self.params
        a_shift = self.gate.add(ctx, E(a), C(1 << p.S))
        div, rem = self.range.div_mod(ctx, E(a_shift), 1 << p.precision_bits, p.A)
        q = self.gate.sub(ctx, E(div), C(1 << (p.S - p.precision_bits)))
        return q, rem

    def qsqrt(self, ctx, a: AssignedValue) -> AssignedValue:
        # parity unpinned (A.6): a single witness cell holding floor(sqrt(a*2^P))
        return ctx.load_witness(isqrt_fixed(a.value, self.P))


# --- the reference API, restated (src/matrix/mod.rs) -------------------------

class ZkVector:
    def __init__(self, v: List[AssignedValue]) -> None:
        self.v = v

    @classmethod
    def new(cls, ctx, fp: FixedPointChip, v: Sequence[float]) -> "ZkVector":  # :29
        return cls([ctx.load_witness(fp.quantization(x)) for x in v])

    def size(self) -> int:
        return len(self.v)

    def dequantize(self, fp) -> List[float]:  # :50
        return [fp.dequantization(e.value) for e in self.v]

    def inner_product(self, ctx, fp, x: List[AssignedValue]) -> AssignedValue:  # :79
        assert self.size() == len(x)
        res_s = fp.gate.inner_product(ctx, [E(e) for e in x], [E(e) for e in self.v])  # :100
        res, _ = fp.signed_div_scale(ctx, res_s)  # :104
        return res

    def _norm_square(self, ctx, fp):  # :111
        return self.inner_product(ctx, fp, self.v)

    def norm(self, ctx, fp):  # :124
        return fp.qsqrt(ctx, self._norm_square(ctx, fp))

    def _dist_square(self, ctx, fp, x):  # :136
        assert self.size() == len(x)
        diff = [fp.qsub(ctx, r, s) for r, s in zip(self.v, x)]
        return ZkVector(diff)._norm_square(ctx, fp)

    def dist(self, ctx, fp, x):  # :156
        return fp.qsqrt(ctx, self._dist_square(ctx, fp, x))

    def mul(self, ctx, fp, a: "ZkMatrix") -> "ZkVector":  # :169
        assert a.num_col == self.size()
        return ZkVector([self.inner_product(ctx, fp, row) for row in a.matrix])

    def entries_less_than(self, ctx, fp, max_bits: int) -> None:  # :185
        for e in self.v:
            fp.range.range_check(ctx, e, max_bits)

    def entries_in_desc_order(self, ctx, fp, max_bits: int) -> None:  # :199
        diffs = [fp.qsub(ctx, self.v[i], self.v[i + 1]) for i in range(len(self.v) - 1)]
        for e in diffs:
            fp.range.range_check(ctx, e, max_bits)


class ZkMatrix:
    def __init__(self, matrix: List[List[AssignedValue]]) -> None:
        self.matrix = matrix
        self.num_rows = len(matrix)
        self.num_col = len(matrix[0])

    @classmethod
    def new(cls, ctx, fp, matrix: Sequence[Sequence[float]]) -> "ZkMatrix":  # :230
        ncol = len(matrix[0])
        for row in matrix:
            assert len(row) == ncol
        return cls([[ctx.load_witness(fp.quantization(x)) for x in row] for row in matrix])

    def dequantize(self, fp):  # :257
        return [[fp.dequantization(e.value) for e in row] for row in self.matrix]

    @staticmethod
    def verify_mul(ctx, fp, a: "ZkMatrix", b: "ZkMatrix", c_s, init_rand: AssignedValue) -> None:  # :299
        assert a.num_col == b.num_rows
        assert len(c_s) == a.num_rows
        assert len(c_s[0]) == b.num_col
        assert len(c_s[0]) >= 1
        d = len(c_s[0])
        one = ctx.load_witness(1)  # :318
        ctx.constants.append((one.index, 1))  # :319 gate.assert_is_const(ctx, &one, &F::ONE): a constant equality, no new cell
        v = [one]
        for i in range(1, d):
            v.append(fp.gate.mul(ctx, E(v[i - 1]), E(init_rand)))  # :324
        cs_v = field_mat_vec_mul(ctx, fp.gate, c_s, v)
        b_v = field_mat_vec_mul(ctx, fp.gate, b.matrix, v)
        ab_v = field_mat_vec_mul(ctx, fp.gate, a.matrix, b_v)
        for x, y in zip(cs_v, ab_v):
            fp.gate.is_equal(ctx, E(x), E(y))  # :340 (result discarded upstream)

    @staticmethod
    def rescale_matrix(ctx, fp, c_s) -> "ZkMatrix":  # :354
        return ZkMatrix([[fp.signed_div_scale(ctx, e)[0] for e in row] for row in c_s])

    @staticmethod
    def transpose_matrix(a: "ZkMatrix") -> "ZkMatrix":  # :408
        return ZkMatrix([[a.matrix[j][i] for j in range(a.num_rows)] for i in range(a.num_col)])


def check_abs_less_than(ctx, rng: RangeChip, x: AssignedValue, bnd: int) -> None:  # :425
    t = rng.gate.add(ctx, E(x), C(bnd - 1))
    rng.check_big_less_than_safe(ctx, t, 2 * bnd - 1)


def check_mat_diff(ctx, rng, a, b, tol: int) -> None:  # :441
    assert len(a) == len(b) and len(a[0]) == len(b[0])
    for i in range(len(a)):
        for j in range(len(a[0])):
            diff = rng.gate.sub(ctx, E(a[i][j]), E(b[i][j]))
            check_abs_less_than(ctx, rng, diff, tol)


def check_mat_id(ctx, rng, a, scalar_id: AssignedValue, tol: int) -> None:  # :461
    zero = ctx.load_constant(0)
    b = [[scalar_id if i == j else zero for j in range(len(a[0]))] for i in range(len(a))]
    check_mat_diff(ctx, rng, a, b, tol)


def check_mat_entries_bounded(ctx, rng, a, bnd: int) -> None:  # :490
    for row in a:
        for e in row:
            check_abs_less_than(ctx, rng, e, bnd)


def honest_prover_mat_mul(ctx, a, b):  # :546
    c_s = field_mat_mul([[e.value for e in r] for r in a], [[e.value for e in r] for r in b])
    return [[ctx.load_witness(e) for e in row] for row in c_s]


def field_mat_vec_mul(ctx, gate: GateChip, a, v):  # :574
    assert len(a[0]) == len(v)
    return [gate.inner_product(ctx, [E(x) for x in row], [E(x) for x in v]) for row in a]


def mat_times_diag_mat(ctx, gate, a, v):  # :610
    assert len(v) <= len(a[0])
    return [[gate.mul(ctx, E(a[i][j]), E(v[j])) for j in range(len(v))] for i in range(len(a))]


def err_calc(p: int, size: int, max_norm: float, eps_svd: float, eps_u: float):
    """reference src/svd/mod.rs:155-163."""
    precision = 2.0 ** (-1.0 * (p + 1.0))
    err_svd = (precision * size * (1.0 + max_norm + eps_svd + precision)
               + size * max_norm * precision
               + (1.0 + eps_u) ** 0.5 * (max_norm + eps_svd) * eps_u
               + (1.0 + eps_u) ** 0.5 * eps_svd)
    err_u = eps_u + precision * size * (2.0 * (1.0 + eps_u) + precision)
    return err_svd, err_u


def check_svd_phase0(ctx, fp, m, u, v, d, err_svd, err_u, max_bits_d):
    """reference src/svd/mod.rs:32-116."""
    P = fp.P
    N, M = m.num_rows, m.num_col
    assert N == u.num_rows and M == v.num_rows
    min_nm = min(N, M)
    assert u.num_rows == u.num_col and v.num_rows == v.num_col and min_nm == len(d.v)
    rng, gate = fp.range, fp.gate
    max_bits = max_bits_d + P
    d.entries_less_than(ctx, fp, max_bits)
    d.entries_in_desc_order(ctx, fp, max_bits)
    unit_bnd_q = (1 << P) + 1
    check_mat_entries_bounded(ctx, rng, u.matrix, unit_bnd_q)
    check_mat_entries_bounded(ctx, rng, v.matrix, unit_bnd_q)
    u_t, v_t = ZkMatrix.transpose_matrix(u), ZkMatrix.transpose_matrix(v)
    if min_nm == M:
        u_times_d = mat_times_diag_mat(ctx, gate, u.matrix, d.v)
    else:
        zero = ctx.load_constant(0)
        u_times_d = mat_times_diag_mat(ctx, gate, u.matrix, d.v)
        for row in u_times_d:
            row.extend([zero] * (M - N))
    m_times_vt = honest_prover_mat_mul(ctx, m.matrix, v_t.matrix)
    err_svd_scale = int(round(err_svd * float(1 << (2 * P))))
    err_u_scale = int(round(err_u * float(1 << (2 * P))))
    check_mat_diff(ctx, rng, u_times_d, m_times_vt, err_svd_scale)
    quant_square = ctx.load_constant((1 << P) * (1 << P))
    u_times_ut = honest_prover_mat_mul(ctx, u.matrix, u_t.matrix)
    check_mat_id(ctx, rng, u_times_ut, quant_square, err_u_scale)
    v_times_vt = honest_prover_mat_mul(ctx, v.matrix, v_t.matrix)
    check_mat_id(ctx, rng, v_times_vt, quant_square, err_u_scale)
    return u_t, v_t, m_times_vt, u_times_ut, v_times_vt


def check_svd_phase1(ctx, fp, m, u, v, u_t, v_t, m_times_vt, u_times_ut, v_times_vt, init_rand):
    """reference src/svd/mod.rs:127-144."""
    ZkMatrix.verify_mul(ctx, fp, m, v_t, m_times_vt, init_rand)
    ZkMatrix.verify_mul(ctx, fp, u, u_t, u_times_ut, init_rand)
    ZkMatrix.verify_mul(ctx, fp, v, v_t, v_times_vt, init_rand)


# --- MockProver-equivalent check of one Context ------------------------------

def mock_prove(contexts, lookup_bits: int) -> List[str]:
    """MockProver-equivalent check of one Context or a list of Contexts (phase 0, phase 1, ...).
    Returns the list of violated constraints (empty == satisfied): gate q*(a + b*c - d) on every
    selected row, copy constraints (also across contexts), constant equalities and lookup-range
    membership (value < 2^lookup_bits)."""
    if isinstance(contexts, Context):
        contexts = [contexts]
    by_id = {c.ctx_id: c for c in contexts}
    errs = []
    for ctx in contexts:
        adv, cid = ctx.advice, ctx.ctx_id
        for i, sel in enumerate(ctx.selector):
            if sel:
                a, b, c, d = adv[i:i + 4]
                if (a + b * c - d) % R_MOD != 0:
                    errs.append(f"gate@{cid}:{i}")
        for (ca, ia), (cb, ib) in ctx.copies:
            if by_id[ca].advice[ia] != by_id[cb].advice[ib]:
                errs.append(f"copy@{ca}:{ia},{cb}:{ib}")
        for i, v in ctx.constants:
            if adv[i] != v:
                errs.append(f"const@{cid}:{i}")
        for i in ctx.lookups:
            if adv[i] >= (1 << lookup_bits):
                errs.append(f"lookup@{cid}:{i}")
    return errs


# --- seeded re-implementation of input-creator.py's distribution -------------

def make_svd_inputs(n: int, m: int, seed: int):
    """input-creator.py:23-52 with a seeded generator (the original is unseeded)."""
    rng = np.random.default_rng(seed)
    mat = rng.uniform(-10.0, 10.0, size=(n, m))
    mat = mat / np.linalg.norm(mat, ord=2) * rng.uniform(1, 100)
    U, D, V = np.linalg.svd(mat)
    wrong = mat.copy()
    wrong[rng.integers(n), rng.integers(m)] += 1e-7
    return dict(m=mat, u=U, d=D, v=V), dict(m=wrong, u=U, d=D, v=V)
