"""the bulk hand-off.  The library's cell layouts + value expansion (host functions, no GPU) are diffed
cell by cell -- values, kinds, gate selectors, lookup cells (in push order), copy constraints, constants -- against the
oracle's halo2-base Context model (oracle/pyoracle.py) running the reference's per-cell code path."""
import importlib
import random

import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po

pkg = importlib.import_module("halo2-svd041_b200")
KIND = {0: "W", 1: "C", 2: "E"}


def _replay(layout, units, first_cell, input_cells):
    """What a binding does with a layout: kinds, selector offsets, lookups, copies, constants of `units` units appended at
    `first_cell`; input_cells[u][i] = advice index of input i of unit u."""
    kinds, gates, lookups, copies, consts = [], [], [], [], []

    def cell(u, off):
        return first_cell + u * layout.cells + off if off >= 0 else input_cells[u][-1 - off]

    cvals = po.unpack_mont(layout.constants) if len(layout.constants) else []
    for u in range(units):
        kinds += [KIND[k] for k in layout.kind]
        gates += [cell(u, g) for g in layout.gates]
        lookups += [cell(u, c) for c in layout.lookups]
        for c, (k, s) in enumerate(zip(layout.kind, layout.source)):
            if k == layout.EXISTING:
                copies.append((cell(u, s), cell(u, c)))
            elif k == layout.CONSTANT:
                consts.append((cell(u, c), cvals[s]))
        copies += [(cell(u, a), cell(u, b)) for a, b in layout.copies]
    return kinds, gates, lookups, copies, consts


def _compare(ctx, start, layout, units, input_cells, values):
    assert len(ctx.advice) - start == units * layout.cells
    assert po.unpack_mont(values.reshape(-1, 4)) == ctx.advice[start:]
    kinds, gates, lookups, copies, consts = _replay(layout, units, start, input_cells)
    assert kinds == ctx.kind[start:]
    assert sorted(gates) == [i for i in range(start, len(ctx.advice)) if ctx.selector[i]]
    assert lookups == ctx.lookups                                     # push order matters
    assert sorted(copies) == sorted((a[1], b[1]) for a, b in ctx.copies)
    assert sorted(consts) == sorted(ctx.constants)


def _field_values(rng, count, P):
    lim = 1 << (2 * P + 10)
    vals = [rng.randrange(-lim, lim) % po.R_MOD for _ in range(count - 4)]
    return vals + [0, 1, po.R_MOD - 1, (1 << (3 * P)) - 1]


@pytest.mark.parametrize("P,lb,S,A", [(32, 19, -1, -1), (42, 19, -1, -1), (63, 19, -1, -1), (32, 12, -1, -1),
                                      (63, 19, 189, 190), (20, 32, -1, -1), (63, 8, -1, -1)])
def test_rescale_cells_match_the_per_cell_path(P, lb, S, A):
    rng = random.Random(P * 100 + lb)
    vals = _field_values(rng, 40, P)
    fp = po.FixedPointChip(P, lb, S, A)
    ctx = po.Context()
    cells = [ctx.load_witness(v) for v in vals]
    start = len(ctx.advice)
    for c in cells:
        fp.signed_div_scale(ctx, c)                    # the reference's per-element path (src/matrix/mod.rs:369)
    cs = po.pack_mont(vals)
    _q, _rem, wit = corac.rescale_witness(cs, P, lb, S, A)
    layout = pkg.CellsLayout.rescale(P, lb, S, A)
    assert layout.witnesses == wit.shape[1] and layout.inputs == 1 and layout.cells == fp.params.cells
    values = layout.expand(cs, wit, len(vals), threads=3)
    _compare(ctx, start, layout, len(vals), [[c.index] for c in cells], values)
    assert (values == layout.expand(cs, wit, len(vals), threads=1)).all()


@pytest.mark.parametrize("bnd,lb,with_diff", [((1 << 42) + 1, 19, True), (12345678901234567890123, 19, False), (5, 19, True),
                                              ((1 << 63) + 1, 12, False)])
def test_abs_less_than_cells_match_the_per_cell_path(bnd, lb, with_diff):
    rng = random.Random(lb + bnd % 1000)
    n = 25
    xs = [rng.randrange(-(bnd - 1), bnd) % po.R_MOD for _ in range(n)]
    ys = [rng.randrange(po.R_MOD) for _ in range(n)]
    rc = po.RangeChip(lb)
    ctx = po.Context()
    if with_diff:
        xs = [(x + y) % po.R_MOD for x, y in zip(xs, ys)]
    xc = [ctx.load_witness(v) for v in xs]
    yc = [ctx.load_witness(v) for v in ys]
    start = len(ctx.advice)
    for a, b in zip(xc, yc):
        if with_diff:                                   # check_mat_diff (:441-459)
            a = rc.gate.sub(ctx, po.E(a), po.E(b))
        po.check_abs_less_than(ctx, rc, a, bnd)         # :425-437
    x, y = po.pack_mont(xs), po.pack_mont(ys)
    wit = corac.abs_less_than_witness(x, bnd, lb, y=y if with_diff else None)
    layout = pkg.CellsLayout.abs_less_than(bnd, lb, with_diff)
    assert layout.witnesses == wit.shape[1] and layout.inputs == (2 if with_diff else 1)
    inputs = np.ascontiguousarray(np.stack([x, y], axis=1)) if with_diff else x
    values = layout.expand(inputs, wit, n)
    _compare(ctx, start, layout, n, [[a.index, b.index] for a, b in zip(xc, yc)], values)


@pytest.mark.parametrize("bits,lb", [(72, 19), (19, 19), (20, 19), (57, 19), (5, 19), (38, 19), (33, 32)])
def test_range_check_cells_match_the_per_cell_path(bits, lb):
    rng = random.Random(bits)
    xs = [rng.randrange(1 << bits) for _ in range(20)] + [0, (1 << bits) - 1]
    rc = po.RangeChip(lb)
    ctx = po.Context()
    xc = [ctx.load_witness(v) for v in xs]
    start = len(ctx.advice)
    for c in xc:
        rc.range_check(ctx, c, bits)                    # ZkVector::entries_less_than (:185-197)
    x = po.pack_mont(xs)
    wit = corac.range_check_witness(x, bits, lb)
    layout = pkg.CellsLayout.range_check(bits, lb)
    assert layout.witnesses == wit.shape[1]
    values = layout.expand(x, wit if wit.shape[1] else None, len(xs))
    _compare(ctx, start, layout, len(xs), [[c.index] for c in xc], values)


def test_verify_mul_cell_stream_matches_the_per_cell_path():
    """The whole advice stream of ZkMatrix::verify_mul (:299-342) assembled from the Witness arrays with the bulk
    expanders, against the oracle running the reference's call order."""
    rng = np.random.default_rng(3)
    n, k, m, P = 5, 4, 6, 42
    a = corac.quantize(rng.uniform(-3, 3, size=(n, k)), P)
    b = corac.quantize(rng.uniform(-3, 3, size=(k, m)), P)
    c = corac.field_mat_mul(a, b)
    c[2, 3, 0] ^= np.uint64(1)                          # one dishonest row: diff != 0, is_zero = 0, inv = diff^-1
    gamma = po.pack_mont([0x1234567890ABCDEF1234567])
    fw = corac.freivalds_witness(a, b, c, gamma)
    # oracle: the reference's code path, cell by cell
    ctx = po.Context()
    load = lambda mat: [[ctx.load_witness(v) for v in po.unpack_mont(row)] for row in mat]   # noqa: E731
    ac, bc, cc = load(a), load(b), load(c)
    g = ctx.load_witness(po.unpack_mont(gamma)[0])
    start = len(ctx.advice)
    po.ZkMatrix.verify_mul(ctx, po.FixedPointChip(P, 19), po.ZkMatrix(ac), po.ZkMatrix(bc), cc, g)
    want = ctx.advice[start:]
    # bulk path
    csv, bv, abv = fw["prefix_cv"][:, -1], fw["prefix_bv"][:, -1], fw["prefix_abv"][:, -1]
    parts = [pkg.expand_gamma_power_cells(gamma, fw["powers"]),
             pkg.expand_inner_product_cells(c, fw["powers"], fw["prefix_cv"]),
             pkg.expand_inner_product_cells(b, fw["powers"], fw["prefix_bv"]),
             pkg.expand_inner_product_cells(a, np.ascontiguousarray(bv), fw["prefix_abv"]),
             pkg.expand_is_equal_cells(np.ascontiguousarray(csv), np.ascontiguousarray(abv), fw["diff"], fw["is_zero"],
                                       fw["inv"])]
    got = [v for p in parts for v in po.unpack_mont(p.reshape(-1, 4))]
    assert got == want
    # the generic layout form of is_equal gives the same cells from an interleaved witness stripe
    layout = pkg.CellsLayout.is_equal()
    assert (layout.cells, layout.witnesses, layout.inputs) == (12, 3, 2)
    stripe = np.ascontiguousarray(np.stack([fw["diff"], fw["is_zero"], fw["inv"]], axis=1))
    inputs = np.ascontiguousarray(np.stack([csv, abv], axis=1))
    assert (layout.expand(inputs, stripe, n) == parts[4]).all()


def test_expand_inner_product_per_row_vector():
    """ZkVector::inner_product (:79-100): u = x, v = self, a vector per row."""
    rng = np.random.default_rng(4)
    x = corac.quantize(rng.uniform(-3, 3, size=(3, 7)), 32)
    s = corac.quantize(rng.uniform(-3, 3, size=(3, 7)), 32)
    pre = corac.zkvec_inner_prefix(x, s)
    got = pkg.expand_inner_product_cells(x, s, pre, per_row_v=True)
    gate = po.GateChip()
    for r in range(3):
        ctx = po.Context()
        xc = [ctx.load_witness(v) for v in po.unpack_mont(x[r])]
        sc = [ctx.load_witness(v) for v in po.unpack_mont(s[r])]
        start = len(ctx.advice)
        gate.inner_product(ctx, [po.E(c) for c in xc], [po.E(c) for c in sc])
        assert po.unpack_mont(got[r]) == ctx.advice[start:]
