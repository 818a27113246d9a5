"""GPU parity tests (run with -m gpu on a B200): every CUDA kernel, called through the C ABI, must
be BIT-EXACT against the CPU oracle on the same seeded inputs, plus size-independent properties at
the BASELINE sizes (Freivalds identity, q*2^P + rem == a + 2^S, limb recomposition, linearity)."""
import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po
from tests.util import adversarial_fr, quantized_matrix, random_fr, raw_limbs

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return a.shape == b.shape and bool((a == b).all())


# ---------------------------------------------------------------- K1: field mat-mul
@pytest.mark.parametrize("n,k,m", [(8, 8, 8), (1, 1, 1), (5, 7, 3), (33, 17, 47), (32, 16, 32), (64, 100, 31),
                                   (256, 256, 256)])
def test_fr_matmul_matches_oracle(handle, n, k, m):
    rng = np.random.default_rng(n * 10007 + k * 101 + m)
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    c = handle.fr_matmul(a, b)
    assert _eq(c, corac.field_mat_mul(a, b, threads=0))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_fr_matmul_all_tile_variants(handle, pkg, variant):
    rng = np.random.default_rng(variant)
    a, b = random_fr(rng, 70, 50, ), random_fr(rng, 50, 45)
    try:
        handle.tune("matmul_karatsuba", 0)      # tile variants of the schoolbook engine
        handle.tune("matmul_variant", variant)
        c = handle.fr_matmul(a, b)
    finally:
        handle.tune("matmul_variant", 0)
        handle.tune("matmul_karatsuba", -1)
    assert _eq(c, corac.field_mat_mul(a, b, threads=0))


@pytest.mark.parametrize("n,k,m", [(8, 8, 8), (33, 100, 47), (128, 1024, 256), (5, 300, 3), (64, 17, 31)])
def test_fr_matmul_streamk_schedule(handle, pkg, n, k, m):
    """Stream-K schedule (flattened (tile, k-chunk) space cut evenly over the resident CTAs + partial-tile
    fix-up) forced on: same bytes as the oracle and as the one-CTA-per-tile schedule."""
    rng = np.random.default_rng(n * 7 + k)
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    try:
        handle.tune("matmul_karatsuba", 0)      # schoolbook engine: this test is about the schedule
        handle.tune("matmul_streamk", 1)
        got = handle.fr_matmul(a, b)
        handle.tune("matmul_streamk", 0)
        plain = handle.fr_matmul(a, b)
    finally:
        handle.tune("matmul_streamk", -1)
        handle.tune("matmul_karatsuba", -1)
    assert _eq(got, plain)
    if n * k * m <= 1 << 22:
        assert _eq(got, corac.field_mat_mul(a, b))
    else:
        rows = [0, n // 2, n - 1]
        assert _eq(got[rows], corac.field_mat_mul(np.ascontiguousarray(a[rows]), b))


@pytest.mark.parametrize("kara", [1, 2, 3])
@pytest.mark.parametrize("streamk", [0, 1])
@pytest.mark.parametrize("n,k,m", [(8, 8, 8), (33, 100, 47), (5, 300, 3), (64, 17, 31), (128, 256, 100)])
def test_fr_matmul_karatsuba_engine(handle, pkg, kara, streamk, n, k, m):
    """Karatsuba engine (operands pre-split at bit 127, three lazy 4x4-limb accumulators, recombination in the
    epilogue), one-CTA-per-tile and stream-K schedules: same bytes as the oracle, incl. adversarial operands."""
    rng = np.random.default_rng(n * 11 + k)
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    adv = adversarial_fr()
    a.reshape(-1, 4)[: min(len(adv), n * k)] = adv[: n * k]
    b.reshape(-1, 4)[-min(len(adv), k * m):] = adv[: min(len(adv), k * m)]
    try:
        handle.tune("matmul_karatsuba", kara)
        handle.tune("matmul_streamk", streamk)
        got = handle.fr_matmul(a, b)
    finally:
        handle.tune("matmul_karatsuba", -1)
        handle.tune("matmul_streamk", -1)
    assert _eq(got, corac.field_mat_mul(a, b))


@pytest.mark.parametrize("n,k,m", [(8, 8, 8), (1, 1, 1), (5, 7, 3), (33, 100, 47), (64, 17, 31), (130, 300, 20),
                                   (128, 1024, 256), (200, 2100, 9), (129, 1025, 17), (256, 256, 256)])
def test_fr_matmul_tensor_core_engine(handle, pkg, n, k, m):
    """Tensor-core engine (tcgen05 kind::i8 over the 32 byte planes of each operand, the 63 diagonal sums
    overlap-added in TMEM, one Montgomery reduction per element) forced on for every shape, incl. ragged tiles,
    k tails, k > 1024 (several accumulation passes) and adversarial operands: same bytes as the oracle."""
    rng = np.random.default_rng(n * 13 + k)
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    adv = adversarial_fr()
    a.reshape(-1, 4)[: min(len(adv), n * k)] = adv[: n * k]
    b.reshape(-1, 4)[-min(len(adv), k * m):] = adv[: min(len(adv), k * m)]
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        handle.tune("matmul_tc", 0)
        imad = handle.fr_matmul(a, b)
    finally:
        handle.tune("matmul_tc", -1)
    assert _eq(got, imad)
    if n * k * m <= 1 << 22:
        assert _eq(got, corac.field_mat_mul(a, b))
    else:
        rows = [0, n // 2, n - 1]
        assert _eq(got[rows], corac.field_mat_mul(np.ascontiguousarray(a[rows]), b))


def test_fr_matmul_tensor_core_worst_case_accumulation(handle, pkg):
    """Every product (r-1)^2 and every byte of one operand 0xff-heavy, k = 4096 (four accumulation passes):
    the per-diagonal sums stay below 2^31 by construction (32 * 1024 * 255^2), so this must be exact."""
    try:
        handle.tune("matmul_tc", 1)
        big = np.ascontiguousarray(np.broadcast_to(raw_limbs([po.R_MOD - 1])[0], (3, 4096, 4)))
        bigt = np.ascontiguousarray(np.broadcast_to(raw_limbs([po.R_MOD - 1])[0], (4096, 5, 4)))
        c = handle.fr_matmul(big, bigt)
        exp = 4096 * (po.R_MOD - 1) ** 2 * po.MONT_RINV % po.R_MOD
        assert [int(x) for x in c[0, 0]] == [int(x) for x in raw_limbs([exp])[0]]
        assert _eq(c, np.broadcast_to(c[0, 0], c.shape))
        # the largest canonical value whose low 31 bytes are all 0xff
        ff = (0x2f << 248) | ((1 << 248) - 1)
        x = np.ascontiguousarray(np.broadcast_to(raw_limbs([ff])[0], (2, 1024, 4)))
        y = np.ascontiguousarray(np.broadcast_to(raw_limbs([ff])[0], (1024, 2, 4)))
        c = handle.fr_matmul(x, y)
        exp = 1024 * ff * ff * po.MONT_RINV % po.R_MOD
        assert [int(v) for v in c[1, 1]] == [int(v) for v in raw_limbs([exp])[0]]
    finally:
        handle.tune("matmul_tc", -1)


def test_fr_matmul_adversarial_operands(handle):
    adv = adversarial_fr()  # 0, 1, r-1, R, R^2, 2^253, ...
    k = adv.shape[0]
    a = np.stack([adv, adv[::-1], np.roll(adv, 3, axis=0)])          # 3 x k
    b = np.stack([adv, np.roll(adv, 5, axis=0)], axis=1)             # k x 2
    assert _eq(handle.fr_matmul(np.ascontiguousarray(a), np.ascontiguousarray(b)),
               corac.field_mat_mul(np.ascontiguousarray(a), np.ascontiguousarray(b)))
    # worst case for the lazy accumulator: every product is (r-1)^2, long k
    big = np.ascontiguousarray(np.broadcast_to(raw_limbs([po.R_MOD - 1])[0], (2, 4096, 4)))
    bigt = np.ascontiguousarray(np.broadcast_to(raw_limbs([po.R_MOD - 1])[0], (4096, 2, 4)))
    c = handle.fr_matmul(big, bigt)
    exp = 4096 * (po.R_MOD - 1) ** 2 * po.MONT_RINV % po.R_MOD
    assert [int(x) for x in c[0, 0]] == [int(x) for x in raw_limbs([exp])[0]]
    assert _eq(c, np.broadcast_to(c[0, 0], c.shape))


def test_fr_matmul_transposed_b(handle):
    rng = np.random.default_rng(11)
    a, bt = random_fr(rng, 19, 23), random_fr(rng, 29, 23)  # b^T is m x k
    b = np.ascontiguousarray(bt.transpose(1, 0, 2))
    assert _eq(handle.fr_matmul(a, bt, b_transposed=True), corac.field_mat_mul(a, b))


def test_fr_matmul_quantized_inputs(handle):
    """input-creator.py distribution, P=63 (BASELINE configs[3] arithmetic at a size the oracle finishes)."""
    rng = np.random.default_rng(3)
    a, b = quantized_matrix(rng, 96, 96, 63), quantized_matrix(rng, 96, 96, 63)
    assert _eq(handle.fr_matmul(a, b), corac.field_mat_mul(a, b, threads=0))


def test_fr_matmul_rejects_bad_arguments(handle, pkg):
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):
        handle.fr_matmul(random_fr(rng, 2, 3), random_fr(rng, 4, 2))   # reference :515 assert
    bad = random_fr(rng, 2, 2)
    bad[0, 0] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)    # >= r
    with pytest.raises(pkg.H2svdError) as ei:
        handle.fr_matmul(bad, random_fr(rng, 2, 2))
    assert ei.value.code == pkg._ffi.ERANGE


def test_fr_matmul_n1024_properties(handle):
    """BASELINE size (N=1024): linearity / sampled rows vs the oracle (a full oracle run takes minutes)."""
    rng = np.random.default_rng(1024)
    N = 1024
    a, b = quantized_matrix(rng, N, N, 63), quantized_matrix(rng, N, N, 63)
    c = handle.fr_matmul(a, b)
    rows = [0, 1, 511, 1023]
    for r in rows:
        assert _eq(c[r:r + 1], corac.field_mat_mul(np.ascontiguousarray(a[r:r + 1]), b, threads=0))
    # Freivalds identity with the oracle's arithmetic: (C v) == A (B v)
    g = random_fr(rng, 1)
    fw = corac.freivalds_witness(a, b, c, g, threads=0)
    assert not np.any(fw["diff"])


# ---------------------------------------------------------------- K2/K3: Freivalds
@pytest.mark.parametrize("n,k,m", [(8, 8, 8), (1, 1, 1), (5, 7, 3), (3, 40, 33), (70, 65, 100), (256, 256, 256)])
def test_freivalds_witness_matches_oracle(handle, n, k, m):
    rng = np.random.default_rng(n + 31 * k + 977 * m)
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    cs = corac.field_mat_mul(a, b, threads=0)
    g = random_fr(rng, 1)
    got = handle.freivalds_witness(a, b, cs, g)
    exp = corac.freivalds_witness(a, b, cs, g, threads=0)
    for key in exp:
        assert _eq(got[key], exp[key]), key
    assert not np.any(got["diff"])


def test_freivalds_witness_dishonest_product(handle):
    rng = np.random.default_rng(5)
    n, k, m = 6, 5, 7
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    cs = corac.field_mat_mul(a, b)
    cs[3, 2] = random_fr(rng, 1)[0]
    g = random_fr(rng, 1)
    got = handle.freivalds_witness(a, b, cs, g)
    exp = corac.freivalds_witness(a, b, cs, g)
    for key in exp:
        assert _eq(got[key], exp[key]), key
    assert np.any(got["diff"][3]) and not np.any(got["is_zero"][3])


def test_gamma_powers_edge_cases(handle):
    ones_a = po.pack_mont([1] * 5).reshape(1, 5, 4)
    ones_b = po.pack_mont([1] * 45).reshape(5, 9, 4)
    cs = corac.field_mat_mul(ones_a, ones_b)
    for g in ([0], [1], [po.R_MOD - 1], [2]):
        gm = po.pack_mont(g)
        got = handle.freivalds_witness(ones_a, ones_b, cs, gm)
        assert _eq(got["powers"], corac.gamma_powers(gm, 9))
        assert not np.any(got["diff"])


# ---------------------------------------------------------------- K4: rescale
@pytest.fixture(params=[0, 1], ids=["sums=step+add", "sums=SmallSum"])
def running_sums(handle, request):
    """Both formulations of the range-check running sums (tuning switch rescale_fast_sums, fr::SmallSum when 1)."""
    handle.tune("rescale_fast_sums", request.param)
    yield request.param
    handle.tune("rescale_fast_sums", 0)


@pytest.mark.parametrize("P,lb,S,A", [(32, 19, -1, -1), (42, 19, -1, -1), (63, 19, -1, -1), (32, 12, -1, -1),
                                      (63, 8, -1, -1), (32, 20, -1, -1), (63, 19, 189, 190), (32, 19, 100, 110)])
def test_rescale_witness_matches_oracle(handle, running_sums, P, lb, S, A):
    rng = np.random.default_rng(P * 1000 + lb)
    pyr = __import__("random").Random(P + lb)
    Sv = 3 * P if S < 0 else S
    vals = [pyr.randrange(-(1 << (Sv - 1)), 1 << (Sv - 1)) % po.R_MOD for _ in range(300)]
    vals += [0, 1, po.R_MOD - 1, (1 << Sv) - 1, (po.R_MOD - (1 << Sv)) % po.R_MOD]
    cs = np.concatenate([po.pack_mont(vals), random_fr(rng, 64), adversarial_fr()])
    q, wit = handle.rescale_witness(cs, P, lb, S, A)
    eq, erem, ewit = corac.rescale_witness(cs, P, lb, S, A)
    assert wit.shape[1] == handle.rescale_witness_count(P, lb, S, A)
    assert _eq(q, eq) and _eq(wit, ewit)


def test_rescale_of_matmul_output_properties(handle):
    """mat-mul -> rescale at P=63 (configs[3] pipeline, reduced N): q*2^P + rem == c + 2^S and limbs
    recompose, checked with Python ints on the GPU output alone."""
    rng = np.random.default_rng(7)
    P, lb, N = 63, 19, 64
    a, b = quantized_matrix(rng, N, N, P), quantized_matrix(rng, N, N, P)
    cs = handle.fr_matmul(a, b)
    q, wit = handle.rescale_witness(cs, P, lb)
    prm = po.RescaleParams(P, lb)
    csi, qi = po.unpack_mont(cs), po.unpack_mont(q)
    for e in range(0, N * N, 97):
        w = po.unpack_mont(wit[e])
        a_shift, rem, div = w[0], w[1], w[2]
        assert a_shift == (csi[e] + (1 << prm.S)) % po.R_MOD
        assert div * (1 << P) + rem == a_shift and rem < (1 << P)
        assert qi[e] == (div - (1 << (prm.S - P))) % po.R_MOD == w[-1]
        limbs = [w[3], w[4]] + [w[4 + 2 * i] for i in range(1, prm.n_d - 1)]  # l0, l1, s1, l2, s2, ...
        assert all(l < (1 << lb) for l in limbs)
        assert sum(l << (lb * i) for i, l in enumerate(limbs)) == div
        # dequantized product is close to the float product
    af = np.array([[po.dequantize(x, P) for x in row] for row in
                   np.array(po.unpack_mont(a), dtype=object).reshape(N, N)], dtype=np.float64)
    bf = np.array([[po.dequantize(x, P) for x in row] for row in
                   np.array(po.unpack_mont(b), dtype=object).reshape(N, N)], dtype=np.float64)
    cf = np.array([po.dequantize(x, P) for x in qi]).reshape(N, N)
    assert np.max(np.abs(cf - af @ bf)) < 1e-9


def test_rescale_large_count_chunked(handle):
    """More than one 2^17-element chunk through the double-buffered host path."""
    rng = np.random.default_rng(9)
    P, lb = 32, 19
    count = (1 << 17) + 4099
    cs = corac.quantize(rng.uniform(-1e6, 1e6, size=count), 40)   # signed ~60-bit field elements
    q, wit = handle.rescale_witness(cs, P, lb)
    eq, _, ewit = corac.rescale_witness(cs, P, lb, threads=0)
    assert _eq(q, eq) and _eq(wit, ewit)


# ---------------------------------------------------------------- K5/K6: ZkVector
@pytest.mark.parametrize("batch,ln", [(1, 1), (1, 4), (3, 31), (5, 32), (4, 33), (7, 100), (64, 1024)])
def test_zkvec_inner_prefix_matches_oracle(handle, batch, ln):
    rng = np.random.default_rng(batch * 7919 + ln)
    x, s = random_fr(rng, batch, ln), random_fr(rng, batch, ln)
    assert _eq(handle.zkvec_inner_prefix(x, s), corac.zkvec_inner_prefix(x, s, threads=0))


def test_config1_full_size_zkvector_batch(handle):
    """BASELINE configs[1] at full size: 4096 batched length-1024 vector pairs, P=32 -- inner-product running sums and qsub
    differences bit-exact against the oracle (8 threads on the host)."""
    rng = np.random.default_rng(41)
    x, s = random_fr(rng, 4096, 1024), random_fr(rng, 4096, 1024)
    assert _eq(handle.zkvec_inner_prefix(x, s), corac.zkvec_inner_prefix(x, s, threads=8))
    assert _eq(handle.zkvec_sub(s, x), corac.zkvec_sub(s, x))


def test_zkvec_sub_and_dist_pipeline(handle):
    rng = np.random.default_rng(21)
    P, lb = 32, 19
    x = quantized_matrix(rng, 6, 50, P)
    s = quantized_matrix(rng, 6, 50, P)
    diff = handle.zkvec_sub(s, x)
    assert _eq(diff, corac.zkvec_sub(s, x))
    pre = handle.zkvec_inner_prefix(diff, diff)      # _norm_square of the difference (reference :147-148)
    assert _eq(pre, corac.zkvec_inner_prefix(diff, diff))
    tot = np.ascontiguousarray(pre[:, -1])
    q, wit = handle.rescale_witness(tot, P, lb)
    eq, _, ewit = corac.rescale_witness(tot, P, lb)
    assert _eq(q, eq) and _eq(wit, ewit)
    root = handle.isqrt_fixed(q, P)
    assert _eq(root, corac.isqrt_fixed(q, P))
    # dequantized distance ~ float distance (loose: the model sqrt is exact to 2^-P)
    xf = np.array([po.dequantize(v, P) for v in po.unpack_mont(x)]).reshape(6, 50)
    sf = np.array([po.dequantize(v, P) for v in po.unpack_mont(s)]).reshape(6, 50)
    df = np.array([po.dequantize(v, P) for v in po.unpack_mont(root)])
    assert np.allclose(df, np.linalg.norm(sf - xf, axis=1), atol=1e-6)


def test_reference_test_zkvector_fixture(handle):
    """Inputs of the reference's own smoke driver test_zkvector (src/matrix/test_matrix.rs:51-92):
    matrix[i][j] = i + j/10 (5x4), v1, v2 at P=32 -- inner product / mat-vec running sums + rescale."""
    P, lb = 32, 19
    N, M = 5, 4
    matrix = np.array([[i + j / 10.0 for j in range(M)] for i in range(N)])
    v1 = np.array([(i + (i * i + 1) / 10.0) if i % 2 == 0 else (-i + (i * i + 1) / 10.0) for i in range(M)])
    v2 = np.array([((1.0 + i ** 3) / 10.0) if i % 2 == 0 else (-(1.0 + i ** 3) / 10.0) for i in range(M)])
    qm, q1, q2 = handle.quantize(matrix, P), handle.quantize(v1, P), handle.quantize(v2, P)
    assert _eq(qm, corac.quantize(matrix, P)) and _eq(q1, corac.quantize(v1, P)) and _eq(q2, corac.quantize(v2, P))
    # zkvec1.inner_product(zkvec2.v): u = x = v2, v = self = v1 (reference :100)
    pre = handle.zkvec_inner_prefix(q2.reshape(1, M, 4), q1.reshape(1, M, 4))
    assert _eq(pre, corac.zkvec_inner_prefix(q2.reshape(1, M, 4), q1.reshape(1, M, 4)))
    q, wit = handle.rescale_witness(np.ascontiguousarray(pre[:, -1]), P, lb)
    ip = po.dequantize(po.unpack_mont(q)[0], P)
    assert abs(ip - float(v1 @ v2)) < 1e-8
    # zkvec1.mul(zkmatrix): rows x inner_product(x = row, self = v1)
    selfs = np.ascontiguousarray(np.broadcast_to(q1, (N, M, 4)))
    pre = handle.zkvec_inner_prefix(qm, selfs)
    assert _eq(pre, corac.zkvec_inner_prefix(qm, selfs))
    q, _ = handle.rescale_witness(np.ascontiguousarray(pre[:, -1]), P, lb)
    got = np.array([po.dequantize(v, P) for v in po.unpack_mont(q)])
    assert np.allclose(got, matrix @ v1, atol=1e-8)


def test_quantize_matches_oracle(handle):
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-100, 100, 1000), [0.0, -0.0, 1.5, -1.5, 1e-12, -1e-12, 0.5 / (1 << 32),
                                                        -0.5 / (1 << 32), 99.99999999, -99.99999999]])
    for P in (32, 42, 63):
        assert _eq(handle.quantize(xs, P), corac.quantize(xs, P))
        # ties and near-ties: odd y = |x|*2^P in [2^52, 2^53), 0.49999999999999994, exact halves
        from tests.test_oracle import quantize_tie_cases
        ties = quantize_tie_cases(P)
        assert _eq(handle.quantize(ties, P), corac.quantize(ties, P))


def test_isqrt_matches_oracle(handle):
    vals = [0, 1, 2, 3, 4, 15, 16, 17, (1 << 64) + 5, (1 << 100) + 12345, (1 << 127), (1 << 128) - 1]
    a = po.pack_mont(vals)
    for P in (32, 63):
        assert _eq(handle.isqrt_fixed(a, P), corac.isqrt_fixed(a, P))


# ---------------------------------------------------------------- device-pointer entry points
def test_dev_entry_points_with_torch_tensors(handle):
    import torch
    rng = np.random.default_rng(77)
    n, k, m = 40, 36, 52
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    dev = torch.device("cuda", handle.device)
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    tb = torch.from_numpy(b.view(np.int64)).to(dev)
    tc = torch.empty((n, m, 4), dtype=torch.int64, device=dev)
    tn = torch.empty_like(tc)
    torch.cuda.synchronize()
    handle.fr_matmul_dev(ta, tb, tc)
    handle.fr_matmul_naive_dev(ta, tb, tn)
    handle.sync()
    exp = corac.field_mat_mul(a, b, threads=0)
    assert _eq(tc.cpu().numpy().view(np.uint64), exp)
    assert _eq(tn.cpu().numpy().view(np.uint64), exp)
    before = handle.launch_count
    handle.fr_matmul_dev(ta, tb, tc)
    handle.sync()
    assert handle.launch_count == before + 1


@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("n,k,m,P,lb", [(8, 8, 8, 42, 19), (64, 64, 64, 63, 19), (130, 300, 20, 63, 19),
                                        (100, 1500, 77, 32, 12), (256, 128, 200, 63, 19)])
def test_fr_matmul_rescale_one_call(handle, pkg, fuse, n, k, m, P, lb):
    """h2svd_fr_matmul_rescale_dev: C, quotients and rescale witnesses of the product in one call, as two kernels
    (default) and with the experimental fused tensor-core epilogue: both byte-identical to the oracle."""
    import torch
    rng = np.random.default_rng(n + k + m + P)
    a, b = quantized_matrix(rng, n, k, P), quantized_matrix(rng, k, m, P)
    dev = torch.device("cuda", handle.device)
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    tb = torch.from_numpy(b.view(np.int64)).to(dev)
    W = handle.rescale_witness_count(P, lb)
    tc = torch.full((n, m, 4), -1, dtype=torch.int64, device=dev)
    tq = torch.full((n, m, 4), -1, dtype=torch.int64, device=dev)
    tw = torch.full((n * m, W, 4), -1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    try:
        handle.tune("fuse_rescale", fuse)
        handle.fr_matmul_rescale_dev(ta, tb, tc, P, lb, tq, tw)
        handle.sync()
    finally:
        handle.tune("fuse_rescale", 0)
    c = corac.field_mat_mul(a, b)
    q, _rem, wit = corac.rescale_witness(c.reshape(-1, 4), P, lb, threads=0)
    assert _eq(tc.cpu().numpy().view(np.uint64), c)
    assert _eq(tq.cpu().numpy().view(np.uint64).reshape(-1, 4), q)
    assert _eq(tw.cpu().numpy().view(np.uint64), wit)


# ---------------------------------------------------------------- fused, slab-pipelined sequence
@pytest.mark.parametrize("rows,k,m,P,bv", [(8, 8, 8, 42, None), (37, 20, 45, 63, (3, 11)), (300, 64, 700, 32, None)])
def test_zkmatrix_mul_witness_fused_matches_oracle(handle, pkg, rows, k, m, P, bv):
    """h2svd_zkmatrix_mul_witness (mat-mul -> rescale -> verify_mul, slabs pipelined against the D2H copies,
    pinned output buffers) returns exactly what the three separate oracle functions return."""
    lb = 19
    rng = np.random.default_rng(rows + k)
    a, b = quantized_matrix(rng, rows, k, P), quantized_matrix(rng, k, m, P)
    gamma = random_fr(rng, 1)
    W = handle.rescale_witness_count(P, lb)
    pinned = pkg.PinnedBuffer((rows * m, W, 4))          # the big one in page-locked memory
    res = handle.zkmatrix_mul_witness(a, b, gamma, P, lb, bv_rows=bv, out={"wit": pinned.array})
    c = corac.field_mat_mul(a, b)
    assert _eq(res["c_s"], c)
    eq, _, ewit = corac.rescale_witness(c.reshape(-1, 4), P, lb)
    assert _eq(res["q"].reshape(-1, 4), eq) and _eq(res["wit"], ewit)
    fw = corac.freivalds_witness(a, b, c, gamma)
    r0, r1 = bv if bv is not None else (0, k)
    assert _eq(res["prefix_bv"], fw["prefix_bv"][r0:r1])
    for key in ("powers", "prefix_cv", "prefix_abv", "diff", "is_zero", "inv"):
        assert _eq(res[key], fw[key]), key
    assert not res["diff"].any()


# ---------------------------------------------------------------- SVD-verifier range-check witnesses
@pytest.mark.parametrize("bnd,lb", [((1 << 42) + 1, 19), (12345678901234567890123, 19), (5, 19), ((1 << 63) + 1, 12), (1 << 100, 8)])
def test_abs_less_than_witness_matches_oracle(handle, running_sums, bnd, lb):
    rng = np.random.default_rng(lb)
    import random as _r
    r = _r.Random(bnd % 997)
    n = 3000
    xs = [r.randrange(-(bnd - 1), bnd) % po.R_MOD for _ in range(n - 3)] + [0, bnd - 1, (-(bnd - 1)) % po.R_MOD]
    x = po.pack_mont(xs)
    y = random_fr(rng, n)
    assert _eq(handle.abs_less_than_witness(x, bnd, lb), corac.abs_less_than_witness(x, bnd, lb))
    xy = corac.zkvec_sub(x, corac.zkvec_sub(np.zeros_like(y), y))   # x + y, so that (x + y) - y is in range
    assert _eq(handle.abs_less_than_witness(xy, bnd, lb, y=y), corac.abs_less_than_witness(xy, bnd, lb, y=y))
    # out-of-range inputs (a dishonest prover) still produce the model's cells
    bad = random_fr(rng, 200)
    assert _eq(handle.abs_less_than_witness(bad, bnd, lb), corac.abs_less_than_witness(bad, bnd, lb))


@pytest.mark.parametrize("bits,lb", [(72, 19), (93, 19), (19, 19), (20, 19), (57, 19), (5, 19), (38, 19), (64, 8), (33, 32)])
def test_range_check_witness_matches_oracle(handle, running_sums, bits, lb):
    import random as _r
    r = _r.Random(bits)
    x = po.pack_mont([r.randrange(1 << bits) for _ in range(1500)] + [0, (1 << bits) - 1])
    assert _eq(handle.range_check_witness(x, bits, lb), corac.range_check_witness(x, bits, lb))
    bad = random_fr(np.random.default_rng(bits), 100)
    assert _eq(handle.range_check_witness(bad, bits, lb), corac.range_check_witness(bad, bits, lb))


def test_mat_times_diag_matches_oracle(handle):
    rng = np.random.default_rng(3)
    a, v = random_fr(rng, 37, 50), random_fr(rng, 41)
    assert _eq(handle.mat_times_diag(a, v), corac.mat_times_diag(a, v))
    with pytest.raises(Exception):
        handle.mat_times_diag(random_fr(rng, 3, 4), random_fr(rng, 5))     # reference :616 assert


# ---------------------------------------------------------------- BASELINE configs[4], one rank's share
def test_config4_one_rank_share(handle):
    """configs[4] (4096x2048 . 2048x4096, P=63, lb=19, 8 GPUs): the 512-row slab one of the 8 ranks owns, at full size
    on this GPU through the device-pointer entry points.  Checked by (1) the Freivalds identity on every row (diff == 0,
    computed by kernels independent of the mat-mul), (2) sampled rows of C and (3) sampled rescale witness stripes,
    bit-exact against the oracle."""
    import torch
    P, lb = 63, 19
    rows, k, m = 512, 2048, 4096
    dev = torch.device("cuda", handle.device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    af = (torch.rand((rows, k), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4
    bf = (torch.rand((k, m), dtype=torch.float64, device=dev, generator=gen) - 0.5) * 4

    def fr(*shape):
        return torch.empty(shape + (4,), dtype=torch.int64, device=dev)

    a, b, c = fr(rows, k), fr(k, m), fr(rows, m)
    handle.quantize_dev(af, P, a)
    handle.quantize_dev(bf, P, b)
    handle.fr_matmul_dev(a, b, c)
    gamma = torch.tensor([[0x1234567, 0x89ABCDEF, 0x13579BDF, 0x2468ACE]], dtype=torch.int64, device=dev)
    powers, pcv, pbv, pabv = fr(m), fr(rows, m), fr(k, m), fr(rows, k)
    diff, isz, inv = fr(rows), fr(rows), fr(rows)
    handle.freivalds_witness_dev(a, b, c, gamma, powers, pcv, pbv, pabv, diff, isz, inv)
    W = handle.rescale_witness_count(P, lb)
    q, wit = fr(rows, m), fr(rows * m, W)
    handle.rescale_witness_dev(c, rows * m, P, lb, q, wit)
    handle.sync()
    assert not bool(diff.any().item()), "Freivalds identity violated"
    one = po.pack_mont([1]).view(np.int64)
    assert bool((isz.cpu().numpy() == one).all())
    # sampled rows of C against the oracle
    sel = [0, 255, 511]
    a_h = a[sel].cpu().numpy().view(np.uint64)
    b_h = b.cpu().numpy().view(np.uint64)
    assert _eq(c[sel].cpu().numpy().view(np.uint64), corac.field_mat_mul(np.ascontiguousarray(a_h), b_h, threads=0))
    # sampled witness stripes against the oracle
    idx = torch.tensor([0, 1, 4095, 4096, 1000003, rows * m - 1], device=dev)
    c_s = c.reshape(-1, 4)[idx].cpu().numpy().view(np.uint64)
    eq, _, ewit = corac.rescale_witness(np.ascontiguousarray(c_s), P, lb)
    assert _eq(wit[idx].cpu().numpy().view(np.uint64), ewit)
    assert _eq(q.reshape(-1, 4)[idx].cpu().numpy().view(np.uint64), eq)
