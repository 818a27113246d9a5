"""Parity against the REFERENCE ITSELF, when its advice dumps are available.

rust/reference-golden/dump_advice.rs (to be run inside the reference tree by anyone with cargo; unverified here -- no Rust
toolchain in this container) writes every advice cell of two of the reference's own circuits as to_repr() hex.  When those
files sit in tests/golden/, these tests diff the oracle's advice stream -- and, on the GPU box, the CUDA path's -- against
them cell by cell.  While they are absent the tests SKIP with the reason "parity unpinned": mat-mul, gamma powers and
running sums are pinned by the field definition alone, signed_div_scale's constants / qsqrt / quantization rounding are not
(SURVEY.md 8c, A.5, A.6)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")
ZKVEC = os.path.join(GOLDEN, "reference_advice_zkvector.hex")
MATMUL = os.path.join(GOLDEN, "reference_advice_matmul_8x8.hex")
GAMMA = 0x123456789ABCDEF0


def _read_hex(path):
    """one cell per line: 32 bytes little-endian canonical (halo2curves to_repr)"""
    with open(path) as fh:
        return [int.from_bytes(bytes.fromhex(ln.strip()), "little") for ln in fh if ln.strip()]


def _zkvector_fixture():
    N, M = 5, 4
    matrix = [[i + j / 10.0 for j in range(M)] for i in range(N)]                         # src/matrix/test_matrix.rs:51-60
    v1 = [(i + (i * i + 1) / 10.0) if i % 2 == 0 else (-i + (i * i + 1) / 10.0) for i in range(M)]   # :65-73
    v2 = [((1.0 + i ** 3) / 10.0) if i % 2 == 0 else (-(1.0 + i ** 3) / 10.0) for i in range(M)]     # :80-88
    return matrix, v1, v2


def oracle_zkvector_stream(lookup_bits=19):
    matrix, v1, v2 = _zkvector_fixture()
    fp = po.FixedPointChip(32, lookup_bits)
    ctx = po.Context()
    zm = po.ZkMatrix.new(ctx, fp, matrix)
    z1 = po.ZkVector.new(ctx, fp, v1)
    z2 = po.ZkVector.new(ctx, fp, v2)
    z1.inner_product(ctx, fp, z2.v)
    z1.norm(ctx, fp)
    z1.dist(ctx, fp, z2.v)
    z1.mul(ctx, fp, zm)
    return ctx.advice


def oracle_matmul_stream(lookup_bits=19):
    with open(os.path.join(GOLDEN, "matrix_8x8.in")) as fh:
        inp = json.load(fh)
    fp = po.FixedPointChip(42, lookup_bits)
    ctx = po.Context()
    a = po.ZkMatrix.new(ctx, fp, inp["m"])
    b = po.ZkMatrix.new(ctx, fp, inp["u"])
    c_s = po.honest_prover_mat_mul(ctx, a.matrix, b.matrix)
    po.ZkMatrix.rescale_matrix(ctx, fp, c_s)
    g = ctx.load_witness(GAMMA)
    po.ZkMatrix.verify_mul(ctx, fp, a, b, c_s, g)
    return ctx.advice


def _diff(name, got, want):
    assert len(got) == len(want), f"{name}: {len(got)} advice cells, the reference has {len(want)}"
    bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
    assert not bad, f"{name}: {len(bad)} cells differ from the reference, first at index {bad[0]}"


def test_oracle_streams_are_well_formed_without_the_reference():
    """Runs always: the two circuits the dump example builds exist in the oracle and have the expected sizes."""
    zk = oracle_zkvector_stream()
    mm = oracle_matmul_stream()
    assert len(zk) > 5 * 4 + 8 and len(mm) > 2 * 64 + 64
    assert all(0 <= v < po.R_MOD for v in zk + mm)


@pytest.mark.skipif(not os.path.exists(ZKVEC), reason="parity unpinned: tests/golden/reference_advice_zkvector.hex not provided "
                                                       "(run rust/reference-golden/dump_advice.rs in the reference tree)")
def test_oracle_matches_reference_zkvector_advice():
    _diff("test_zkvector circuit (oracle)", oracle_zkvector_stream(), _read_hex(ZKVEC))


@pytest.mark.skipif(not os.path.exists(MATMUL), reason="parity unpinned: tests/golden/reference_advice_matmul_8x8.hex not provided "
                                                        "(run rust/reference-golden/dump_advice.rs in the reference tree)")
def test_oracle_matches_reference_matmul_advice():
    _diff("8x8 mat-mul + rescale + verify_mul circuit (oracle)", oracle_matmul_stream(), _read_hex(MATMUL))


def _cuda_matmul_stream(pkg):
    with open(os.path.join(GOLDEN, "matrix_8x8.in")) as fh:
        inp = json.load(fh)
    P, lb = 42, 19
    gamma = po.pack_mont([GAMMA])
    with pkg.Handle() as h:
        a = h.quantize(np.array(inp["m"]), P)
        b = h.quantize(np.array(inp["u"]), P)
        c = h.fr_matmul(a, b)
        q, wit = h.rescale_witness(c, P, lb)
        fw = h.freivalds_witness(a, b, c, gamma)
    layout = pkg.CellsLayout.rescale(P, lb)
    cells = [po.unpack_mont(a.reshape(-1, 4)), po.unpack_mont(b.reshape(-1, 4)), po.unpack_mont(c.reshape(-1, 4)),
             po.unpack_mont(layout.expand(np.ascontiguousarray(c.reshape(-1, 4)), wit, c.shape[0] * c.shape[1]).reshape(-1, 4)),
             [GAMMA],
             po.unpack_mont(pkg.expand_gamma_power_cells(gamma, fw["powers"])),
             po.unpack_mont(pkg.expand_inner_product_cells(c, fw["powers"], fw["prefix_cv"]).reshape(-1, 4)),
             po.unpack_mont(pkg.expand_inner_product_cells(b, fw["powers"], fw["prefix_bv"]).reshape(-1, 4)),
             po.unpack_mont(pkg.expand_inner_product_cells(a, np.ascontiguousarray(fw["prefix_bv"][:, -1]), fw["prefix_abv"]).reshape(-1, 4)),
             po.unpack_mont(pkg.expand_is_equal_cells(np.ascontiguousarray(fw["prefix_cv"][:, -1]),
                                                      np.ascontiguousarray(fw["prefix_abv"][:, -1]), fw["diff"], fw["is_zero"],
                                                      fw["inv"]).reshape(-1, 4))]
    return [v for part in cells for v in part]


@pytest.mark.gpu
def test_cuda_full_advice_stream_matches_oracle(pkg):
    """The CUDA path's COMPLETE advice stream (Witness, Existing and Constant cells, assembled with the bulk expanders) of
    ZkMatrix::new x2 -> honest_prover_mat_mul -> rescale_matrix -> verify_mul on the committed 8x8 input is the oracle's."""
    assert _cuda_matmul_stream(pkg) == oracle_matmul_stream(19)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(MATMUL), reason="parity unpinned: reference advice dump not provided")
def test_cuda_path_matches_reference_matmul_advice(pkg):
    _diff("8x8 mat-mul + rescale + verify_mul circuit (CUDA)", _cuda_matmul_stream(pkg), _read_hex(MATMUL))
