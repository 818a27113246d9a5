"""Small ragged-size pass over every kernel for compute-sanitizer (developer tool, run under gpurun):
    python tools/sanitize_small.py && compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("halo2-svd041_b200")
from oracle import corac  # noqa: E402  (checker only)


def rfr(rng, *shape):
    a = rng.integers(0, 1 << 64, size=shape + (4,), dtype=np.uint64)
    a[..., 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    rng = np.random.default_rng(0)
    h = pkg.Handle(0)
    for (n, k, m) in ((5, 7, 3), (33, 100, 47), (70, 65, 100), (128, 200, 130)):
        a, b = rfr(rng, n, k), rfr(rng, k, m)
        for kara in (0, 3):
            for sk in (0, 1):
                h.tune("matmul_karatsuba", kara)
                h.tune("matmul_streamk", sk)
                c = h.fr_matmul(a, b)
                assert (c == corac.field_mat_mul(a, b)).all(), (n, k, m, kara, sk)
        h.tune("matmul_karatsuba", -1)
        h.tune("matmul_streamk", -1)
        g = rfr(rng, 1)
        fw = h.freivalds_witness(a, b, c, g)
        ew = corac.freivalds_witness(a, b, c, g)
        assert all((fw[key] == ew[key]).all() for key in ew)
        for P, lb in ((63, 19), (32, 12)):
            q, w = h.rescale_witness(c, P, lb)
            eq, _, ewit = corac.rescale_witness(c.reshape(-1, 4), P, lb)
            assert (w == ewit).all() and (q.reshape(-1, 4) == eq).all()
        res = h.zkmatrix_mul_witness(a, b, g, 42, 19, bv_rows=(1, k - 1))
        assert (res["c_s"] == c).all()
    # tensor-core engines on small quantized operands: every tile width of the small-operand engine, with and without
    # clusters; the full-width engine; both running-sum formulations of the rescale; the one-call step
    from tests.util import quantized_matrix  # noqa: E402
    n, k, m = 130, 200, 60
    a, b = quantized_matrix(rng, n, k, 63), quantized_matrix(rng, k, m, 63)
    want = corac.field_mat_mul(a, b)
    h.tune("matmul_tc", 1)
    for width in (8, 16, 24, 28):
        for cluster in (0, 2):
            h.tune("matmul_small_width", width)
            h.tune("matmul_cluster", cluster)
            assert (h.fr_matmul(a, b) == want).all(), (width, cluster)
            assert h.last_matmul_engine() == "tensor-small"
    h.tune("matmul_small_width", 0)
    h.tune("matmul_cluster", 0)
    h.tune("matmul_small", 0)
    assert (h.fr_matmul(a, b) == want).all() and h.last_matmul_engine() == "tensor"
    h.tune("matmul_small", -1)
    h.tune("matmul_tc", -1)
    for fast in (1, 0):
        h.tune("rescale_fast_sums", fast)
        q, w = h.rescale_witness(want, 63, 19)
        eq, _, ewit = corac.rescale_witness(want.reshape(-1, 4), 63, 19)
        assert (w == ewit).all() and (q.reshape(-1, 4) == eq).all()
    res = h.zkmatrix_mul_witness(a, b, rfr(rng, 1), 63, 19)
    assert (res["c_s"] == want).all() and not res["diff"].any()
    for batch, ln in ((3, 31), (700, 130), (5, 1000), (600, 129)):
        x, s = rfr(rng, batch, ln), rfr(rng, batch, ln)
        assert (h.zkvec_inner_prefix(x, s) == corac.zkvec_inner_prefix(x, s)).all()
        assert (h.zkvec_sub(s, x) == corac.zkvec_sub(s, x)).all()
    x = rfr(rng, 777)
    assert (h.abs_less_than_witness(x, (1 << 42) + 1, 19) == corac.abs_less_than_witness(x, (1 << 42) + 1, 19)).all()
    assert (h.abs_less_than_witness(x, 12345, 19, y=x[::-1].copy()) == corac.abs_less_than_witness(x, 12345, 19, y=x[::-1].copy())).all()
    assert (h.range_check_witness(x, 93, 19) == corac.range_check_witness(x, 93, 19)).all()
    assert (h.mat_times_diag(rfr(rng, 9, 11), x[:10]) is not None)
    q = h.quantize(rng.uniform(-5, 5, size=(37, 5)), 42)
    assert (h.isqrt_fixed(h.quantize(rng.uniform(0, 5, size=50), 32), 32) is not None) and q is not None
    h.close()
    print("sanitize_small: ok")


if __name__ == "__main__":
    main()
