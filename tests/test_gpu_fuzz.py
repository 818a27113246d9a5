"""Seeded differential fuzzing of the CUDA library against the C oracle: random shapes around the tile / dispatch
boundaries of every kernel, random chip parameters, full-width random and adversarial operands.  Bit-exact or fail."""
import random

import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po
from tests.util import adversarial_fr, random_fr

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return a.shape == b.shape and bool((a == b).all())


def _sprinkle(rng, arr):
    """overwrite a few random positions with adversarial field elements (0, 1, r-1, R, ...)"""
    adv = adversarial_fr()
    flat = arr.reshape(-1, 4)
    for _ in range(min(8, flat.shape[0])):
        flat[rng.integers(flat.shape[0])] = adv[rng.integers(len(adv))]
    return arr


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_matmul_shapes_engines_schedules(handle, pkg, seed):
    r = random.Random(1000 + seed)
    rng = np.random.default_rng(seed)
    edges = [1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 100, 127, 128, 129, 200]
    for _ in range(6):
        n, k, m = r.choice(edges), r.choice(edges + [256, 300]), r.choice(edges)
        a, b = _sprinkle(rng, random_fr(rng, n, k)), _sprinkle(rng, random_fr(rng, k, m))
        want = corac.field_mat_mul(a, b, threads=8)
        try:
            for kara in (-1, 0, 1, 3):
                for sk in (-1, 0, 1):
                    handle.tune("matmul_karatsuba", kara)
                    handle.tune("matmul_streamk", sk)
                    assert _eq(handle.fr_matmul(a, b), want), (n, k, m, kara, sk)
            handle.tune("matmul_karatsuba", -1)
            handle.tune("matmul_streamk", -1)
            handle.tune("matmul_tc", 1)             # tensor-core engine forced, whatever the shape
            assert _eq(handle.fr_matmul(a, b), want), (n, k, m, "tensor-core")
            handle.tune("matmul_tc", -1)
            bt = np.ascontiguousarray(b.transpose(1, 0, 2))
            assert _eq(handle.fr_matmul(a, bt, b_transposed=True), want), (n, k, m, "transposed")
        finally:
            handle.tune("matmul_karatsuba", -1)
            handle.tune("matmul_streamk", -1)
            handle.tune("matmul_tc", -1)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_mat_vec_prefix_dispatch_boundaries(handle, seed):
    """row counts around the one-warp-per-row / row-splitting dispatch threshold (4 x SM count) and lengths around the
    128-element tile and the 32-lane chunk"""
    r = random.Random(2000 + seed)
    rng = np.random.default_rng(100 + seed)
    sm4 = 4 * handle.sm_count
    for _ in range(5):
        rows = r.choice([1, 2, 7, sm4 - 1, sm4, sm4 + 1, 700])
        ln = r.choice([1, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 513, 1000])
        x, s = _sprinkle(rng, random_fr(rng, rows, ln)), _sprinkle(rng, random_fr(rng, rows, ln))
        assert _eq(handle.zkvec_inner_prefix(x, s), corac.zkvec_inner_prefix(x, s, threads=8)), (rows, ln)
    # shared vector (verify_mul shape) through the Freivalds entry point, ragged
    n, k, m = r.choice([3, 40, 129]), r.choice([5, 130, 257]), r.choice([2, 128, 300])
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    c = corac.field_mat_mul(a, b, threads=8)
    g = random_fr(rng, 1)
    fw, ew = handle.freivalds_witness(a, b, c, g), corac.freivalds_witness(a, b, c, g)
    assert all(_eq(fw[key], ew[key]) for key in ew), (n, k, m)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_rescale_parameters(handle, seed):
    r = random.Random(3000 + seed)
    rng = np.random.default_rng(200 + seed)
    for _ in range(5):
        P = r.randint(1, 63)
        lb = r.randint(8, 32)
        S = r.choice([-1, r.randint(P, min(252, 3 * P + 5))])
        A = r.choice([-1, r.randint(P + 1, min(250, 4 * P))])
        if corac.rescale_witness_count(P, lb, S, A) < 0:
            continue
        count = r.choice([1, 31, 32, 33, 127, 129, 1000])
        x = _sprinkle(rng, random_fr(rng, count))
        q, wit = handle.rescale_witness(x, P, lb, S, A)
        eq, _, ewit = corac.rescale_witness(x, P, lb, S, A)
        assert _eq(q.reshape(-1, 4), eq) and _eq(wit, ewit), (P, lb, S, A, count)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_range_check_helpers(handle, seed):
    r = random.Random(4000 + seed)
    rng = np.random.default_rng(300 + seed)
    for _ in range(6):
        lb = r.randint(8, 32)
        count = r.choice([1, 33, 500])
        x = _sprinkle(rng, random_fr(rng, count))
        bits = r.randint(1, min(250, lb * 30))
        assert _eq(handle.range_check_witness(x, bits, lb), corac.range_check_witness(x, bits, lb)), (bits, lb)
        bnd = r.getrandbits(r.randint(1, 200)) + 1
        if corac.lib().orc_abs_less_than_witness_count(corac._p(corac._bnd(bnd)), lb, 0) < 0:
            continue
        y = random_fr(rng, count)
        assert _eq(handle.abs_less_than_witness(x, bnd, lb), corac.abs_less_than_witness(x, bnd, lb)), (bnd, lb)
        assert _eq(handle.abs_less_than_witness(x, bnd, lb, y=y), corac.abs_less_than_witness(x, bnd, lb, y=y)), (bnd, lb)
