"""The C++ host mirror of the reference API (include/h2svd_zk.hpp: Context / GateChip / RangeChip /
FixedPointChip041 / ZkVector / ZkMatrix with the reference's names and call order) against the oracle's
model of the reference (oracle/pyoracle.py): every advice cell, cell kind, gate selector, copy constraint,
constant and lookup must be identical, and the recorded constraint system must be satisfied.

Two builds of the same mirror + the same scenarios (tests/host/zk_host_shim.cpp):
  * CPU (not gpu): linked against tests/host/abi_over_oracle.c -- checks the host bookkeeping only;
  * GPU (-m gpu):  linked against the product, libh2svd_b200.so -- values come from the CUDA kernels.
The scenarios are the reference's own drivers: test_zkvector (src/matrix/test_matrix.rs:39-198),
test_field_mat_times_vec (:201-265) and the README.md:34-47 ZkMatrix sequence (the missing test_zkmatrix)."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

from oracle import corac
from oracle import pyoracle as po
from tests.util import ROOT

HOST = os.path.join(ROOT, "tests", "host")
KIND = {"W": 0, "E": 1, "C": 2}


def _build(tmp, gpu: bool) -> ct.CDLL:
    so = os.path.join(tmp, "libzkh_gpu.so" if gpu else "libzkh_cpu.so")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HOST, "zk_host_shim.cpp")]
    if gpu:
        pkg_dir = os.path.join(ROOT, "halo2-svd041_b200")
        cmd += ["-DZKH_WITH_EXPAND", os.path.join(pkg_dir, "libh2svd_b200.so"), f"-Wl,-rpath,{pkg_dir}"]
    else:
        corac.build()
        obj = os.path.join(tmp, "abi_over_oracle.o")
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-c", os.path.join(HOST, "abi_over_oracle.c"), "-o", obj])
        odir = os.path.join(ROOT, "oracle")
        cmd += [obj, os.path.join(odir, "libfr_oracle.so"), f"-Wl,-rpath,{odir}"]
    subprocess.check_call(cmd)
    lib = ct.CDLL(so)
    for name in ("zkh_ctx_count", "zkh_ctx_len", "zkh_ctx_ncopies", "zkh_ctx_nconstants", "zkh_ctx_nlookups", "zkh_nscalars"):
        getattr(lib, name).restype = ct.c_size_t
        getattr(lib, name).argtypes = [] if name in ("zkh_ctx_count", "zkh_nscalars") else [ct.c_size_t]
    lib.zkh_error.restype = ct.c_char_p
    lib.zkh_failure.restype = ct.c_char_p
    lib.zkh_failure.argtypes = [ct.c_int]
    lib.zkh_run_zkmatrix.argtypes = [ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_size_t, ct.c_size_t, ct.c_size_t, ct.c_void_p]
    lib.zkh_run_zkvector.argtypes = [ct.c_int]
    if gpu:
        lib.zkh_run_rescale_bulk.argtypes = [ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_size_t, ct.c_size_t, ct.c_size_t]
    lib.zkh_run_svd.argtypes = [ct.c_int, ct.c_int] + [ct.c_void_p] * 4 + [ct.c_size_t] * 3 + [ct.c_void_p]
    lib.zkh_run_mat_times_vec.argtypes = [ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_size_t, ct.c_size_t]
    lib.zkh_run_bad_shapes.argtypes = [ct.c_int]
    lib.zkh_run_dishonest_product.argtypes = [ct.c_int, ct.c_int]
    lib.zkh_ctx_export.argtypes = [ct.c_size_t] + [ct.c_void_p] * 7
    lib.zkh_scalars.argtypes = [ct.c_void_p]
    return lib


@pytest.fixture(scope="module")
def cpu_lib(tmp_path_factory):
    return _build(str(tmp_path_factory.mktemp("zkh_cpu")), gpu=False)


@pytest.fixture(scope="module")
def gpu_lib(tmp_path_factory):
    return _build(str(tmp_path_factory.mktemp("zkh_gpu")), gpu=True)


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def _export(lib):
    out = []
    for c in range(lib.zkh_ctx_count()):
        n, nc, nk, nl = lib.zkh_ctx_len(c), lib.zkh_ctx_ncopies(c), lib.zkh_ctx_nconstants(c), lib.zkh_ctx_nlookups(c)
        adv = np.zeros((n, 4), dtype=np.uint64)
        kind, sel = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
        copies = np.zeros((nc, 4), dtype=np.uint64)
        kidx, kval = np.zeros(nk, dtype=np.uint64), np.zeros((nk, 4), dtype=np.uint64)
        look = np.zeros(nl, dtype=np.uint64)
        lib.zkh_ctx_export(c, _p(adv), _p(kind), _p(sel), _p(copies), _p(kidx), _p(kval), _p(look))
        out.append(dict(advice=adv, kind=kind, selector=sel, copies=copies, const_idx=kidx, const_val=kval, lookups=look))
    return out


def _scalars(lib):
    s = np.zeros(lib.zkh_nscalars(), dtype=np.float64)
    if s.size:
        lib.zkh_scalars(_p(s))
    return s


def _assert_same_context(got: dict, ctx: po.Context):
    """Identical constraint system: every advice value, kind, selector, copy, constant, lookup."""
    assert len(ctx.advice) == got["advice"].shape[0]
    want = po.pack_mont(ctx.advice)
    bad = np.nonzero((want != got["advice"]).any(axis=1))[0]
    assert bad.size == 0, f"first differing advice cell: {bad[0]}"
    assert [KIND[k] for k in ctx.kind] == got["kind"].tolist()
    assert [int(s) for s in ctx.selector] == got["selector"].tolist()
    assert [[a[0], a[1], b[0], b[1]] for a, b in ctx.copies] == got["copies"].tolist()
    assert [i for i, _ in ctx.constants] == got["const_idx"].tolist()
    assert (po.pack_mont([v for _, v in ctx.constants]).reshape(-1, 4) == got["const_val"]).all()
    assert list(ctx.lookups) == got["lookups"].tolist()


GAMMA = 0x1234567890ABCDEF1234567890ABCDEF0123456789ABCDEF % po.R_MOD


def _oracle_zkmatrix(P, lb, a, b):
    fp = po.FixedPointChip(P, lb)
    ctx, ctx1 = po.Context(0), po.Context(1)
    za, zb = po.ZkMatrix.new(ctx, fp, a.tolist()), po.ZkMatrix.new(ctx, fp, b.tolist())
    c_s = po.honest_prover_mat_mul(ctx, za.matrix, zb.matrix)
    c = po.ZkMatrix.rescale_matrix(ctx, fp, c_s)
    init_rand = ctx1.load_witness(GAMMA)
    po.ZkMatrix.verify_mul(ctx1, fp, za, zb, c_s, init_rand)
    return ctx, ctx1, np.array(c.dequantize(fp)).ravel()


def _check_zkmatrix(lib, P, lb, n, k, m, seed):
    rng = np.random.default_rng(seed)

    def mat(r, c):   # input-creator.py:23-28
        x = rng.uniform(-10.0, 10.0, size=(r, c))
        return np.ascontiguousarray(x / np.linalg.norm(x, ord=2) * rng.uniform(1, 100))

    a, b = mat(n, k), mat(k, m)
    g = po.pack_mont([GAMMA])
    rc = lib.zkh_run_zkmatrix(P, lb, _p(a), _p(b), n, k, m, _p(g))
    assert rc == 0, (lib.zkh_error(), lib.zkh_failure(0))
    ctx, ctx1, c_deq = _oracle_zkmatrix(P, lb, a, b)
    got = _export(lib)
    assert len(got) == 2
    _assert_same_context(got[0], ctx)
    _assert_same_context(got[1], ctx1)
    assert po.mock_prove([ctx, ctx1], lb) == []
    # and the rescaled product is the real-number product to fixed-point accuracy
    assert np.allclose(_scalars(lib), c_deq, rtol=0, atol=0)
    assert np.allclose(c_deq, (a @ b).ravel(), atol=k * 2.0 ** (-P + 8))


def _check_zkvector(lib, lb):
    rc = lib.zkh_run_zkvector(lb)
    assert rc == 0, (lib.zkh_error(), lib.zkh_failure(0))
    P, N, M = 32, 5, 4
    fp = po.FixedPointChip(P, lb)
    ctx = po.Context(0)
    matrix = [[i + j / 10.0 for j in range(M)] for i in range(N)]
    v1 = [(i + (i * i + 1) / 10.0) if i % 2 == 0 else (-i + (i * i + 1) / 10.0) for i in range(M)]
    v2 = [((1.0 + i ** 3) / 10.0) if i % 2 == 0 else (-(1.0 + i ** 3) / 10.0) for i in range(M)]
    zm = po.ZkMatrix.new(ctx, fp, matrix)
    z1, z2 = po.ZkVector.new(ctx, fp, v1), po.ZkVector.new(ctx, fp, v2)
    res = [z1.inner_product(ctx, fp, z2.v), z1.norm(ctx, fp), z2.norm(ctx, fp), z1.dist(ctx, fp, z2.v),
           z1._norm_square(ctx, fp), z2._norm_square(ctx, fp), z1._dist_square(ctx, fp, z2.v)]
    res += z1.mul(ctx, fp, zm).v + z2.mul(ctx, fp, zm).v
    got = _export(lib)
    assert len(got) == 1
    _assert_same_context(got[0], ctx)
    assert po.mock_prove([ctx], lb) == []
    deq = _scalars(lib)
    assert np.array_equal(deq, np.array([fp.dequantization(r.value) for r in res]))
    # the f64 ground truth the reference prints next to the circuit values (test_matrix.rs:99-197)
    a1, a2, mm = np.array(v1), np.array(v2), np.array(matrix)
    truth = [a1 @ a2, np.linalg.norm(a1), np.linalg.norm(a2), np.linalg.norm(a1 - a2), a1 @ a1, a2 @ a2,
             (a1 - a2) @ (a1 - a2)] + list(mm @ a1) + list(mm @ a2)
    assert np.allclose(deq, truth, atol=1e-4)


def _check_mat_times_vec(lib, lb):
    rng = np.random.default_rng(5)
    n, m = 5, 5
    mat = np.ascontiguousarray(rng.uniform(-100.0, 100.0, size=(n, m)))   # test_matrix.rs:216-222
    vec = np.ascontiguousarray(rng.uniform(-100.0, 100.0, size=m))
    rc = lib.zkh_run_mat_times_vec(lb, _p(mat), _p(vec), n, m)
    assert rc == 0, (lib.zkh_error(), lib.zkh_failure(0))
    fp = po.FixedPointChip(32, lb)
    ctx = po.Context(0)
    zm, zv = po.ZkMatrix.new(ctx, fp, mat.tolist()), po.ZkVector.new(ctx, fp, vec.tolist())
    out = [fp.signed_div_scale(ctx, x)[0] for x in po.field_mat_vec_mul(ctx, fp.gate, zm.matrix, zv.v)]
    _assert_same_context(_export(lib)[0], ctx)
    assert np.allclose(_scalars(lib), mat @ vec, atol=1e-5)
    assert np.array_equal(_scalars(lib), np.array([fp.dequantization(o.value) for o in out]))


def _oracle_svd(inputs, P, lb, err_size):
    fp = po.FixedPointChip(P, lb)
    ctx = po.Context(0)
    m, u, v = (po.ZkMatrix.new(ctx, fp, inputs[k]) for k in ("m", "u", "v"))
    d = po.ZkVector.new(ctx, fp, inputs["d"])
    err_svd, err_u = po.err_calc(P, err_size, 100.0, 1e-10, 1e-10)
    out = po.check_svd_phase0(ctx, fp, m, u, v, d, err_svd, err_u, 30)
    ctx1 = po.Context(1)
    po.check_svd_phase1(ctx1, fp, m, u, v, *out, ctx1.load_witness(GAMMA))
    return ctx, ctx1


def _run_svd(lib, inputs, P, lb, err_size):
    arrs = [np.ascontiguousarray(np.array(inputs[k], dtype=np.float64)) for k in ("m", "u", "v", "d")]
    n, mm = arrs[0].shape
    g = po.pack_mont([GAMMA])
    return lib.zkh_run_svd(P, lb, *[_p(a) for a in arrs], n, mm, err_size, _p(g))


def _check_svd(lib, n, mm, P, lb):
    """BASELINE configs[0]: the SVD circuit (phase 0 + phase 1) on an input-creator.py style matrix: every cell of
    both contexts identical to the oracle's model, `matrix` satisfied, `matrix-wrong` rejected by a phase-0 check."""
    good, wrong = po.make_svd_inputs(n, mm, 2024 + n)
    good = {k: np.asarray(v).tolist() for k, v in good.items()}
    wrong = {k: np.asarray(v).tolist() for k, v in wrong.items()}
    rc = _run_svd(lib, good, P, lb, max(n, mm))
    assert rc == 0, (lib.zkh_error(), lib.zkh_failure(0))
    ctx, ctx1 = _oracle_svd(good, P, lb, max(n, mm))
    got = _export(lib)
    _assert_same_context(got[0], ctx)
    _assert_same_context(got[1], ctx1)
    assert po.mock_prove([ctx, ctx1], lb) == []
    rc = _run_svd(lib, wrong, P, lb, max(n, mm))
    assert rc >= 0, lib.zkh_error()
    if P >= 42:   # the +1e-7 of input-creator.py:50 is above the tolerance only at the example's precision (P=42) and up
        assert rc > 0, "matrix-wrong must violate the circuit"
    ctxw, ctxw1 = _oracle_svd(wrong, P, lb, max(n, mm))
    gotw = _export(lib)
    _assert_same_context(gotw[0], ctxw)
    _assert_same_context(gotw[1], ctxw1)
    fails = [lib.zkh_failure(i).decode() for i in range(min(rc, 16))]
    assert all(f.startswith("ctx 0:") for f in fails), fails     # rejected in phase 0 (check_mat_diff), never in phase 1
    assert len(po.mock_prove([ctxw, ctxw1], lb)) == rc


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("n,mm,P", [(8, 8, 42), (4, 6, 32), (6, 4, 63)])
def test_host_mirror_svd_circuit_cpu(cpu_lib, n, mm, P):
    _check_svd(cpu_lib, n, mm, P, 19)


@pytest.mark.parametrize("P,lb,n,k,m", [(42, 19, 8, 8, 8), (32, 19, 3, 5, 4), (63, 19, 4, 4, 4), (32, 12, 2, 3, 2),
                                        (32, 32, 2, 2, 3), (63, 8, 2, 3, 2), (42, 25, 3, 2, 2)])
def test_host_mirror_zkmatrix_cpu(cpu_lib, P, lb, n, k, m):
    _check_zkmatrix(cpu_lib, P, lb, n, k, m, seed=100 + P + n)


def test_host_mirror_zkvector_fixture_cpu(cpu_lib):
    _check_zkvector(cpu_lib, 19)


def test_host_mirror_mat_times_vec_cpu(cpu_lib):
    _check_mat_times_vec(cpu_lib, 19)


def test_host_mirror_shape_asserts_cpu(cpu_lib):
    assert cpu_lib.zkh_run_bad_shapes(19) == -1          # the reference panics at src/matrix/mod.rs:515
    assert b"a[0].len() == b.len()" in cpu_lib.zkh_error()


def _check_dishonest(lib):
    """The reference's verify_mul constrains nothing about the Freivalds comparison (is_equal result discarded,
    src/matrix/mod.rs:339-341): a wrong product passes.  The opt-in strict variant rejects it in phase 1."""
    assert lib.zkh_run_dishonest_product(19, 0) == 0, lib.zkh_error()
    rc = lib.zkh_run_dishonest_product(19, 1)
    assert rc == 1 and lib.zkh_failure(0).decode().startswith("ctx 1: constant violated"), (rc, lib.zkh_error())


def test_verify_mul_reference_bug_and_strict_fix_cpu(cpu_lib):
    _check_dishonest(cpu_lib)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_verify_mul_reference_bug_and_strict_fix_gpu(gpu_lib):
    _check_dishonest(gpu_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("P,lb,n,k,m", [(42, 19, 8, 8, 8), (32, 19, 3, 5, 4), (63, 19, 16, 12, 20), (32, 12, 2, 3, 2),
                                        (32, 32, 2, 2, 3), (63, 8, 2, 3, 2), (42, 25, 3, 2, 2)])
def test_host_mirror_zkmatrix_gpu(gpu_lib, P, lb, n, k, m):
    """BASELINE configs[0] shape (8x8, P=42, lb=19) and friends through the CUDA library."""
    _check_zkmatrix(gpu_lib, P, lb, n, k, m, seed=100 + P + n)


@pytest.mark.gpu
@pytest.mark.parametrize("n,mm,P", [(8, 8, 42), (4, 6, 32), (6, 4, 63), (16, 16, 42)])
def test_host_mirror_svd_circuit_gpu(gpu_lib, n, mm, P):
    """configs[0] through the CUDA library: `matrix` passes, `matrix-wrong` fails, every cell as in the oracle."""
    _check_svd(gpu_lib, n, mm, P, 19)


@pytest.mark.gpu
def test_host_mirror_zkvector_fixture_gpu(gpu_lib):
    _check_zkvector(gpu_lib, 19)


@pytest.mark.gpu
def test_host_mirror_mat_times_vec_gpu(gpu_lib):
    _check_mat_times_vec(gpu_lib, 19)


@pytest.mark.gpu
def test_host_mirror_shape_asserts_gpu(gpu_lib):
    assert gpu_lib.zkh_run_bad_shapes(19) == -1
    assert b"a[0].len() == b.len()" in gpu_lib.zkh_error()


@pytest.mark.gpu
@pytest.mark.parametrize("P,lb,n,k,m", [(42, 19, 5, 4, 6), (63, 19, 9, 7, 8), (32, 12, 3, 3, 3)])
def test_bulk_hand_off_leaves_the_same_context_as_the_per_cell_path(gpu_lib, P, lb, n, k, m):
    """SURVEY.md 8(f)3 end to end in the C++ mirror: ZkMatrix::rescale_matrix_bulk (one GPU call, one h2svd_expand_cells pass,
    ONE append of the complete cell stream + a replay of the per-unit layout) and ZkMatrix::rescale_matrix (~100
    assign_region pushes per element) leave byte-identical contexts -- advice, kinds, selectors, copy constraints (same
    order), constants, lookups -- and return the same cells."""
    rng = np.random.default_rng(P + n)
    a = np.ascontiguousarray(rng.uniform(-3, 3, size=(n, k)))
    b = np.ascontiguousarray(rng.uniform(-3, 3, size=(k, m)))
    assert gpu_lib.zkh_run_rescale_bulk(P, lb, _p(a), _p(b), n, k, m) == 0, gpu_lib.zkh_error().decode()
    per_cell, bulk = _export(gpu_lib)
    for key in per_cell:
        assert per_cell[key].shape == bulk[key].shape and (per_cell[key] == bulk[key]).all(), key
    cells = _scalars(gpu_lib)
    assert (cells[: n * m] == cells[n * m:]).all()
