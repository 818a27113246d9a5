"""Shared helpers for tests: seeded inputs in the wire format (numpy uint64[..., 4])."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402

R_LIMBS = np.array([(po.R_MOD >> (64 * i)) & po.MASK64 for i in range(4)], dtype=np.uint64)


def random_fr(rng: np.random.Generator, *shape) -> np.ndarray:
    """Uniform canonical field elements over the WHOLE range [0, r) as raw limbs (any canonical limb pattern is a valid
    Montgomery-form element, so no conversion is needed).  Rejection sampling below 2^254 (r ~ 0.76 * 2^254)."""
    n = int(np.prod(shape)) if shape else 1
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    while True:
        bad = np.zeros(n, dtype=bool)       # bad = (a >= r), compared limb by limb from the top
        undecided = np.ones(n, dtype=bool)
        for limb in (3, 2, 1, 0):
            bad |= undecided & (a[:, limb] > R_LIMBS[limb])
            undecided &= a[:, limb] == R_LIMBS[limb]
        bad |= undecided                    # equal to r
        nb = int(bad.sum())
        if nb == 0:
            break
        fresh = rng.integers(0, 1 << 64, size=(nb, 4), dtype=np.uint64)
        fresh[:, 3] &= np.uint64((1 << 62) - 1)
        a[bad] = fresh
    return a.reshape(shape + (4,))


def adversarial_fr() -> np.ndarray:
    vals = [0, 1, 2, po.R_MOD - 1, po.R_MOD - 2, po.MONT_R, po.MONT_R2, (1 << 253), (po.R_MOD - 1) // 2,
            (1 << 64) - 1, (1 << 128) - 1, (1 << 192) - 1]
    return raw_limbs(vals)


def raw_limbs(vals) -> np.ndarray:
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, x in enumerate(vals):
        for j in range(4):
            out[i, j] = (x >> (64 * j)) & po.MASK64
    return out


def quantized_matrix(rng: np.random.Generator, n: int, m: int, P: int) -> np.ndarray:
    """input-creator.py:23-28 distribution, quantized by the oracle (host, small sizes)."""
    from oracle import corac
    mat = rng.uniform(-10.0, 10.0, size=(n, m))
    mat = mat / np.linalg.norm(mat, ord=2) * rng.uniform(1, 100)
    return corac.quantize(mat, P)
