"""GPU parity tests of the round-2 pieces: the small-operand tensor-core mat-mul engine, the lazily accumulated (b v) totals,
the two-stream device-pointer step, CUDA-graph replay, the single-process multi-GPU handle.  Everything goes through the C
ABI and is compared bit for bit with the CPU oracle."""
import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po
from tests.util import quantized_matrix, random_fr, raw_limbs

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return a.shape == b.shape and bool((a == b).all())


def _signed_matrix(rng, rows, cols, bits):
    """uniform signed integers in (-2^bits, 2^bits) as Montgomery-form field elements (negatives as r - |x|)"""
    import random
    r = random.Random(int(rng.integers(1 << 30)))
    lim = (1 << bits) - 1
    vals = [r.randrange(-lim, lim + 1) for _ in range(rows * cols)]
    return po.pack_mont([v % po.R_MOD for v in vals]).reshape(rows, cols, 4)


# ---------------------------------------------------------------- small-operand engine
@pytest.mark.parametrize("n,k,m", [(64, 64, 64), (96, 80, 72), (128, 128, 24), (129, 130, 25), (1, 1024, 300), (300, 257, 47),
                                   (256, 256, 256), (40, 1000, 23)])
def test_small_operand_engine_matches_oracle_and_full_engine(handle, n, k, m):
    """Quantized (P = 63) operands: the device detects that every element is a small signed integer and runs the 9 x 9
    signed-digit engine; same bytes as the oracle and as the full-width engine, on full, ragged and one-row shapes."""
    rng = np.random.default_rng(n * 31 + k)
    a, b = quantized_matrix(rng, n, k, 63), quantized_matrix(rng, k, m, 63)
    want = corac.field_mat_mul(a, b, threads=0)
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor-small"
        handle.tune("matmul_small", 0)
        full = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor"
    finally:
        handle.tune("matmul_tc", -1)
        handle.tune("matmul_small", -1)
    assert _eq(got, want)
    assert _eq(full, want)


def test_small_operand_engine_extremes_and_signs(handle):
    """+-(2^70 - 1) everywhere (worst case for the 32-bit diagonal accumulators at k = 1024), mixed signs, zeros."""
    lim = (1 << 70) - 1
    n, k, m = 130, 1024, 50
    av = [lim if (i + j) % 3 else -lim for i in range(n) for j in range(k)]
    bv = [lim if (i * 7 + j) % 5 else -lim for i in range(k) for j in range(m)]
    a = po.pack_mont([v % po.R_MOD for v in av]).reshape(n, k, 4)
    b = po.pack_mont([v % po.R_MOD for v in bv]).reshape(k, m, 4)
    a[5, :] = 0
    b[:, 7] = 0
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor-small"
    finally:
        handle.tune("matmul_tc", -1)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))
    assert not got[5].any() and not got[:, 7].any()


def test_small_operand_engine_multi_pass_k(handle):
    """k = 8300 > 8192: two accumulation passes of the small engine (the second adds to the stored first)."""
    rng = np.random.default_rng(77)
    n, k, m = 40, 8300, 30
    a, b = _signed_matrix(rng, n, k, 69), _signed_matrix(rng, k, m, 69)
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor-small"
    finally:
        handle.tune("matmul_tc", -1)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


@pytest.mark.parametrize("where", ["a", "b", "pad"])
def test_small_operand_engine_falls_back_on_one_large_element(handle, where):
    """ONE element at +-2^70 (just out of range) or a full-width element anywhere switches the whole product to the
    full-width engine on the device -- same bytes either way."""
    rng = np.random.default_rng(5)
    n, k, m = 100, 140, 90
    a, b = quantized_matrix(rng, n, k, 63), quantized_matrix(rng, k, m, 63)
    big = po.pack_mont([(1 << 70) % po.R_MOD, (-(1 << 70)) % po.R_MOD])
    if where == "a":
        a[n - 1, k - 1] = big[0]
    elif where == "b":
        b[k - 1, 0] = big[1]
    else:
        a[3, 4] = random_fr(rng, 1)[0]
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor"
    finally:
        handle.tune("matmul_tc", -1)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


def test_full_width_engine_top_byte_planes(handle):
    """Uniform canonical operands over the WHOLE range [0, r) (top byte up to 0x30) plus elements whose every byte is
    large: the highest byte planes of the full-width engine see big values on random inputs too."""
    rng = np.random.default_rng(11)
    n, k, m = 140, 300, 70
    a, b = random_fr(rng, n, k), random_fr(rng, k, m)
    assert int(a[..., 3].max() >> np.uint64(56)) >= 0x2f     # the generator really reaches the top of the range
    hi = raw_limbs([po.R_MOD - 1 - i for i in range(16)])
    a.reshape(-1, 4)[:16] = hi
    b.reshape(-1, 4)[-16:] = hi
    try:
        handle.tune("matmul_tc", 1)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor"
    finally:
        handle.tune("matmul_tc", -1)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


def test_fr_matmul_n1024_full_matrix_both_engines(handle):
    """BASELINE configs[3] size, EVERY element of C: quantized operands (small engine) and uniform full-width operands
    (full engine) against the multi-threaded oracle."""
    rng = np.random.default_rng(1024)
    n = 1024
    a, b = quantized_matrix(rng, n, n, 63), quantized_matrix(rng, n, n, 63)
    got = handle.fr_matmul(a, b)
    assert handle.last_matmul_engine() == "tensor-small"
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))
    a, b = random_fr(rng, n, n), random_fr(rng, n, n)
    got = handle.fr_matmul(a, b)
    assert handle.last_matmul_engine() == "tensor"
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


# ---------------------------------------------------------------- (b v) totals without running sums
@pytest.mark.parametrize("rows,ln", [(1, 1), (3, 31), (5, 33), (64, 1024), (700, 300), (1024, 1024)])
def test_mat_vec_totals_equals_last_running_sum(handle, rows, ln):
    import torch
    rng = np.random.default_rng(rows + ln)
    a, v = random_fr(rng, rows, ln), random_fr(rng, ln)
    dev = torch.device("cuda", handle.device)
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    tv = torch.from_numpy(v.view(np.int64)).to(dev)
    tot = torch.full((rows, 4), -1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    handle.mat_vec_totals_dev(ta, tv, tot)
    handle.sync()
    want = corac.mat_vec_prefix(a, v, threads=0)[:, -1]
    assert _eq(tot.cpu().numpy().view(np.uint64), want)


# ---------------------------------------------------------------- two-stream device step, graph replay
def _dev_step_buffers(torch, dev, rows, k, m, W, bvn):
    def fr(*shape):
        return torch.full(shape + (4,), -1, dtype=torch.int64, device=dev)
    return dict(c_s=fr(rows, m), q=fr(rows, m), wit=fr(rows * m, W), powers=fr(m), prefix_cv=fr(rows, m),
                prefix_bv=fr(max(bvn, 1), m), prefix_abv=fr(rows, k), diff=fr(rows), is_zero=fr(rows), inv=fr(rows))


def _check_step(out, a, b, gamma, P, lb, bv):
    host = {key: t.cpu().numpy().view(np.uint64) for key, t in out.items()}
    c = corac.field_mat_mul(a, b, threads=0)
    assert _eq(host["c_s"], c)
    eq, _, ewit = corac.rescale_witness(c.reshape(-1, 4), P, lb, threads=0)
    assert _eq(host["q"].reshape(-1, 4), eq) and _eq(host["wit"], ewit)
    fw = corac.freivalds_witness(a, b, c, gamma, threads=0)
    r0, r1 = bv
    if r1 > r0:
        assert _eq(host["prefix_bv"][: r1 - r0], fw["prefix_bv"][r0:r1])
    for key in ("powers", "prefix_cv", "prefix_abv", "diff", "is_zero", "inv"):
        assert _eq(host[key], fw[key]), key
    assert not host["diff"].any()


@pytest.mark.parametrize("rows,k,m,P,bv,quant", [(8, 8, 8, 42, None, True), (37, 20, 45, 63, (3, 11), True),
                                                 (130, 200, 150, 63, (50, 50), True), (128, 256, 192, 63, (32, 64), False),
                                                 (300, 64, 500, 32, None, True)])
def test_zkmatrix_mul_witness_dev_and_graph_replay(handle, rows, k, m, P, bv, quant):
    """h2svd_zkmatrix_mul_witness_dev (mat-vecs forked onto the side stream inside the call) is bit-exact with the oracle;
    recorded into a CUDA graph and replayed on fresh inputs it is bit-exact again."""
    import torch
    lb = 19
    rng = np.random.default_rng(rows * 3 + k)
    dev = torch.device("cuda", handle.device)
    W = handle.rescale_witness_count(P, lb)
    r0, r1 = bv if bv is not None else (0, k)
    out = _dev_step_buffers(torch, dev, rows, k, m, W, r1 - r0)
    ta = torch.empty((rows, k, 4), dtype=torch.int64, device=dev)
    tb = torch.empty((k, m, 4), dtype=torch.int64, device=dev)
    tg = torch.empty((1, 4), dtype=torch.int64, device=dev)

    def inputs(seed):
        g = np.random.default_rng(seed)
        if quant:
            return quantized_matrix(g, rows, k, P), quantized_matrix(g, k, m, P), random_fr(g, 1)
        return random_fr(g, rows, k), random_fr(g, k, m), random_fr(g, 1)

    def upload(a, b, gamma):
        ta.copy_(torch.from_numpy(a.view(np.int64)))
        tb.copy_(torch.from_numpy(b.view(np.int64)))
        tg.copy_(torch.from_numpy(gamma.view(np.int64)))
        torch.cuda.synchronize()

    def step():
        handle.zkmatrix_mul_witness_dev(ta, tb, tg, P, lb, bv_rows=(r0, r1), **out)

    a, b, gamma = inputs(int(rng.integers(1 << 30)))
    upload(a, b, gamma)
    step()
    handle.sync()
    _check_step(out, a, b, gamma, P, lb, (r0, r1))
    # record, then replay on new inputs (and once more on yet other inputs)
    launches0 = handle.launch_count
    handle.graph_begin()
    step()
    graph = handle.graph_end()
    assert handle.launch_count == launches0        # recording runs nothing
    try:
        for rep in range(2):
            a, b, gamma = inputs(1000 + rep)
            for t in out.values():
                t.fill_(-1)
            upload(a, b, gamma)
            before = handle.launch_count
            graph.launch()
            handle.sync()
            assert handle.launch_count > before
            _check_step(out, a, b, gamma, P, lb, (r0, r1))
    finally:
        graph.close()


def test_graph_capture_refuses_to_grow_workspaces(pkg):
    """A capture on a cold handle would need cudaMalloc + a stream sync: refused with EINVAL, the handle stays usable."""
    import torch
    with pkg.Handle() as h:
        dev = torch.device("cuda", h.device)
        x = torch.zeros((64, 64, 4), dtype=torch.int64, device=dev)
        c = torch.zeros((64, 64, 4), dtype=torch.int64, device=dev)
        h.graph_begin()
        with pytest.raises(pkg.H2svdError):
            h.fr_matmul_dev(x, x, c)
        try:
            h.graph_end().close()
        except pkg.H2svdError:
            pass          # an invalidated capture may fail to end; either way the handle must work afterwards
        h.fr_matmul_dev(x, x, c)
        h.sync()
        assert not bool(c.any().item())


# ---------------------------------------------------------------- one process, several handles / GPUs
@pytest.mark.parametrize("ndev", [1, 2, 3])
def test_multi_handle_row_sharded_witness_matches_oracle(pkg, ndev):
    """h2svd_multi_zkmatrix_mul_witness: rows of A / C and rows of b . v partitioned over `ndev` handles (on as many GPUs as
    the box has, wrapping round -- the partitioning is the same), assembled output bit-exact with the oracle."""
    import torch
    ngpu = torch.cuda.device_count()
    devices = [i % ngpu for i in range(ndev)]
    rng = np.random.default_rng(ndev)
    n, k, m, P, lb = 70, 45, 130, 63, 19
    a, b = quantized_matrix(rng, n, k, P), quantized_matrix(rng, k, m, P)
    gamma = random_fr(rng, 1)
    with pkg.MultiHandle(devices) as mh:
        assert mh.count == ndev
        res = mh.zkmatrix_mul_witness(a, b, gamma, P, lb)
        assert mh.launch_count() > 0
    c = corac.field_mat_mul(a, b, threads=0)
    assert _eq(res["c_s"], c)
    eq, _, ewit = corac.rescale_witness(c.reshape(-1, 4), P, lb, threads=0)
    assert _eq(res["q"].reshape(-1, 4), eq) and _eq(res["wit"], ewit)
    fw = corac.freivalds_witness(a, b, c, gamma, threads=0)
    for key in ("powers", "prefix_cv", "prefix_bv", "prefix_abv", "diff", "is_zero", "inv"):
        assert _eq(res[key], fw[key]), key


def test_multi_handle_more_handles_than_rows_and_errors(pkg):
    import torch
    ngpu = torch.cuda.device_count()
    rng = np.random.default_rng(9)
    a, b = quantized_matrix(rng, 2, 9, 42), quantized_matrix(rng, 9, 5, 42)
    gamma = random_fr(rng, 1)
    with pkg.MultiHandle([i % ngpu for i in range(4)]) as mh:      # 4 handles, 2 rows: two handles stay idle
        res = mh.zkmatrix_mul_witness(a, b, gamma, 42, 19)
        c = corac.field_mat_mul(a, b)
        assert _eq(res["c_s"], c)
        fw = corac.freivalds_witness(a, b, c, gamma)
        assert _eq(res["prefix_bv"], fw["prefix_bv"]) and _eq(res["prefix_abv"], fw["prefix_abv"])
        bad = a.copy()
        bad[1, 3] = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)    # >= r on the second handle's row
        with pytest.raises(pkg.H2svdError) as ei:
            mh.zkmatrix_mul_witness(bad, b, gamma, 42, 19)
        assert ei.value.code == pkg._ffi.ERANGE
        res2 = mh.zkmatrix_mul_witness(a, b, gamma, 42, 19)       # still usable, no stale flag
        assert _eq(res2["c_s"], c)
    with pytest.raises(pkg.H2svdError):
        pkg.MultiHandle([ngpu + 7])


# ---------------------------------------------------------------- several-warps-per-row mat-vec prefix kernel
@pytest.mark.parametrize("rows,ln", [(1, 256), (3, 300), (5, 512), (7, 1000), (16, 1024), (2, 2100), (3, 4099), (130, 1024)])
@pytest.mark.parametrize("per_row_v", [False, True])
def test_mat_vec_prefix_seg_kernel_matches_oracle(handle, rows, ln, per_row_v):
    """mat_vec_prefix_seg_kernel (SEGS warps per row, tile totals exchanged through shared memory) forced on: every
    running sum bit-exact with the oracle, shared vector (field_mat_vec_mul) and per-row vectors (ZkVector::inner_product),
    ragged last tiles, rows longer than one round of SEGS tiles."""
    import torch
    rng = np.random.default_rng(rows * 7 + ln)
    a = random_fr(rng, rows, ln)
    dev = torch.device("cuda", handle.device)
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    out = torch.full((rows, ln, 4), -1, dtype=torch.int64, device=dev)
    try:
        handle.tune("matvec_seg", 1)
        if per_row_v:
            s = random_fr(rng, rows, ln)
            ts = torch.from_numpy(s.view(np.int64)).to(dev)
            torch.cuda.synchronize()
            handle.zkvec_inner_prefix_dev(ta, ts, out)
            want = corac.zkvec_inner_prefix(a, s, threads=0)
        else:
            v = random_fr(rng, ln)
            tv = torch.from_numpy(v.view(np.int64)).to(dev)
            tot = torch.full((rows, 4), -1, dtype=torch.int64, device=dev)
            torch.cuda.synchronize()
            handle.mat_vec_prefix_dev(ta, tv, out, tot)
            want = corac.mat_vec_prefix(a, v, threads=0)
        handle.sync()
    finally:
        handle.tune("matvec_seg", -1)
    assert _eq(out.cpu().numpy().view(np.uint64), want)
    if not per_row_v:
        assert _eq(tot.cpu().numpy().view(np.uint64), want[:, -1])


def test_freivalds_witness_with_seg_kernel_two_jobs(handle):
    """verify_mul's paired launch (c_s . v and b . v in one grid: two jobs) through the several-warps-per-row kernel."""
    rng = np.random.default_rng(21)
    n, k, m = 40, 70, 600
    a, b = quantized_matrix(rng, n, k, 63), quantized_matrix(rng, k, m, 63)
    c = corac.field_mat_mul(a, b, threads=0)
    gamma = random_fr(rng, 1)
    try:
        handle.tune("matvec_seg", 1)
        fw = handle.freivalds_witness(a, b, c, gamma)
    finally:
        handle.tune("matvec_seg", -1)
    want = corac.freivalds_witness(a, b, c, gamma, threads=0)
    for key, val in want.items():
        assert _eq(fw[key], val), key


@pytest.mark.parametrize("width", [8, 16, 24, 28])
@pytest.mark.parametrize("n,k,m", [(129, 130, 50), (300, 257, 47), (128, 1024, 100), (40, 8300, 30), (256, 256, 57)])
def test_small_operand_engine_every_tile_width(handle, width, n, k, m):
    """The four tile widths of the small-operand engine (128 x 8 / 16 / 24 / 28: MMA N = 80 / 144 / 240 / 252 + 4 zero rows;
    the 28-column tile splits its columns 8 / 8 / 8 / 4 over the epilogue warps) forced in turn on ragged
    shapes (m not a multiple of any width, n not a multiple of 128, k tails, two accumulation passes): same bytes."""
    rng = np.random.default_rng(n + k + m)
    a, b = _signed_matrix(rng, n, k, 69), _signed_matrix(rng, k, m, 69)
    try:
        handle.tune("matmul_tc", 1)
        handle.tune("matmul_small_width", width)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor-small"
    finally:
        handle.tune("matmul_tc", -1)
        handle.tune("matmul_small_width", 0)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


@pytest.mark.parametrize("P,lb,S,A", [(63, 19, -1, -1), (32, 19, -1, -1), (20, 32, -1, -1), (42, 12, -1, -1), (63, 19, 189, 190)])
def test_rescale_store_paths_write_the_same_bytes(handle, P, lb, S, A):
    """The witness stream through per-row bulk copies (0), TMA tensor stores (1) and coalesced 16-byte stores (2): same
    bytes, incl. stripe widths that are not a multiple of the 128-byte TMA box (partial last burst, clipped by the TMA unit)
    and a partial last warp; nothing is written past the end."""
    import torch
    rng = np.random.default_rng(P + lb)
    count = 1000 + 37       # odd: the even class of the TMA path has one row more than the odd class; partial last warp
    cs = random_fr(rng, count)
    cs[:300] = quantized_matrix(rng, 300, 1, P).reshape(300, 4)
    dev = torch.device("cuda", handle.device)
    W = handle.rescale_witness_count(P, lb, S, A)
    tcs = torch.from_numpy(cs.view(np.int64)).to(dev)
    outs = []
    # the store paths (0 = auto: 256-byte-aligned TMA boxes when W % 8 == 4 and the array is 256-byte aligned, else bulk
    # copies; 1 TMA, 2 STG, 3 bulk), also with the witness array starting 1 and 4 witnesses past a 256-byte boundary (falls back
    # to bulk copies), and with the running sums computed both ways (fr::SmallSum / Montgomery step + modular add)
    for store, off, fast in ((0, 0, 0), (1, 0, 1), (2, 0, 1), (3, 0, 0), (3, 0, 1), (0, 1, 1), (0, 4, 0), (0, 8, 0)):
        q = torch.full((count, 4), -1, dtype=torch.int64, device=dev)
        flat = torch.full(((count + 3) * W + 8, 4), -1, dtype=torch.int64, device=dev)   # 3 guard stripes after the end
        wit = flat[off:off + (count + 3) * W].view(count + 3, W, 4)
        torch.cuda.synchronize()
        try:
            handle.tune("rescale_store", store)
            handle.tune("rescale_fast_sums", fast)
            handle.rescale_witness_dev(tcs, count, P, lb, q, wit, S, A)
            handle.sync()
        finally:
            handle.tune("rescale_store", 0)
            handle.tune("rescale_fast_sums", 0)
        assert bool((wit[count:] == -1).all().item()) and bool((flat[:off] == -1).all().item()), "wrote outside the witness array"
        outs.append((q.cpu().numpy().view(np.uint64), wit[:count].cpu().numpy().view(np.uint64)))
    eq, _, ewit = corac.rescale_witness(cs, P, lb, S, A, threads=0)
    for q, wit in outs:
        assert _eq(q, eq) and _eq(wit, ewit)


@pytest.mark.parametrize("x2", [0, 1])
@pytest.mark.parametrize("seg", [0, 1])
@pytest.mark.parametrize("rows,ln", [(3, 300), (700, 131), (16, 1024), (600, 129), (2, 2100)])
def test_mat_vec_prefix_interleaved_products_match_oracle(handle, x2, seg, rows, ln):
    """mont_mul_fast_x2 (two Montgomery products with interleaved carry chains) inside both mat-vec prefix kernels: odd
    tile tails exercise the single-product remainder."""
    import torch
    rng = np.random.default_rng(rows + ln + seg)
    a, v = random_fr(rng, rows, ln), random_fr(rng, ln)
    dev = torch.device("cuda", handle.device)
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    tv = torch.from_numpy(v.view(np.int64)).to(dev)
    out = torch.full((rows, ln, 4), -1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    try:
        handle.tune("matvec_x2", x2)
        handle.tune("matvec_seg", seg)
        handle.mat_vec_prefix_dev(ta, tv, out)
        handle.sync()
    finally:
        handle.tune("matvec_x2", 1)
        handle.tune("matvec_seg", -1)
    assert _eq(out.cpu().numpy().view(np.uint64), corac.mat_vec_prefix(a, v, threads=0))


@pytest.mark.parametrize("cluster", [2])
@pytest.mark.parametrize("width", [8, 16, 24, 28])
@pytest.mark.parametrize("shape", [(1024, 64, 1000), (130, 200, 56), (257, 129, 9), (128, 1024, 24)])
def test_small_operand_engine_cluster_multicast(handle, cluster, width, shape):
    """Clusters of 2 CTAs on the same 128 rows of A: every A-plane stage is loaded once per cluster (each CTA
    multicasts a slice).  Odd numbers of column tiles (surplus CTAs that only feed their peers), partial row blocks and
    several K blocks: the same bytes as the oracle."""
    rng = np.random.default_rng(80 + cluster + width)
    n, k, m = shape
    a, b = _signed_matrix(rng, n, k, 69), _signed_matrix(rng, k, m, 69)
    try:
        handle.tune("matmul_small_width", width)
        handle.tune("matmul_cluster", cluster)
        got = handle.fr_matmul(a, b)
        assert handle.last_matmul_engine() == "tensor-small"
    finally:
        handle.tune("matmul_small_width", 0)
        handle.tune("matmul_cluster", 0)
    assert _eq(got, corac.field_mat_mul(a, b, threads=0))


@pytest.mark.parametrize("rows,k,m", [(300, 200, 260), (128, 1024, 1024)])
def test_device_step_alternate_schedule_writes_the_same_bytes(handle, rows, k, m):
    """tuning switch step_schedule = 1 (all of verify_mul after the mat-mul, through the low-register mat-vec kernel that fits
    next to the rescale CTAs): every output of the one-call step is byte-identical to the default schedule's."""
    import torch
    rng = np.random.default_rng(rows + k)
    P, lb = 63, 19
    a, b = quantized_matrix(rng, rows, k, P), quantized_matrix(rng, k, m, P)
    gamma = random_fr(rng, 1)
    dev = torch.device("cuda", handle.device)
    def t(x):
        return torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).to(dev)
    W = handle.rescale_witness_count(P, lb)
    outs = []
    for sched in (0, 1):
        def z(*shape):
            return torch.full(shape + (4,), -1, dtype=torch.int64, device=dev)
        bufs = dict(c_s=z(rows, m), q=z(rows, m), wit=z(rows * m, W), powers=z(m), prefix_cv=z(rows, m), prefix_bv=z(k, m),
                    prefix_abv=z(rows, k), diff=z(rows), is_zero=z(rows), inv=z(rows))
        try:
            handle.tune("step_schedule", sched)
            handle.zkmatrix_mul_witness_dev(t(a), t(b), t(gamma), P, lb, **bufs)
            handle.sync()
        finally:
            handle.tune("step_schedule", 0)
        outs.append({key: val.cpu().numpy() for key, val in bufs.items()})
    for key in outs[0]:
        assert np.array_equal(outs[0][key], outs[1][key]), key
    assert _eq(outs[1]["c_s"].view(np.uint64), corac.field_mat_mul(a, b, threads=0))
    assert not outs[1]["diff"].any()


def test_graph_is_refused_after_a_workspace_reallocation(pkg):
    """A recorded graph holds workspace pointers: once a larger call has made the handle reallocate, replaying the old graph
    is refused (EINVAL) instead of writing through stale pointers."""
    import torch
    with pkg.Handle() as h:
        dev = torch.device("cuda", h.device)
        def z(*s):
            return torch.zeros(s + (4,), dtype=torch.int64, device=dev)
        a, b, c = z(64, 64), z(64, 64), z(64, 64)
        h.fr_matmul_dev(a, b, c)
        h.sync()
        h.graph_begin()
        h.fr_matmul_dev(a, b, c)
        g = h.graph_end()
        g.launch()
        h.sync()
        a2, b2, c2 = z(256, 256), z(256, 256), z(256, 256)
        h.fr_matmul_dev(a2, b2, c2)          # larger operand planes: the workspace grows
        h.sync()
        with pytest.raises(pkg.H2svdError) as ei:
            g.launch()
        assert ei.value.code == pkg._ffi.EINVAL
        g.close()
