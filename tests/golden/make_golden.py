"""Generates the committed golden fixtures (run HERE, where /root/reference exists; the GPU box only reads
the fixtures).

  1. matrix_8x8.in / matrix-wrong_8x8.in -- produced by the REFERENCE's own input generator,
     /root/reference/input-creator.py, executed unmodified under runpy in a scratch directory with numpy's
     global seed fixed (the script is unseeded; seeding the global generator is the only intervention).
  2. golden.json -- for those inputs (BASELINE configs[0]: P=42, lb=19) and for the reference's test_zkvector
     fixture (src/matrix/test_matrix.rs:51-92, P=32): SHA-256 digests and lengths of every recorded context
     of the oracle's model of the reference (advice stream in Montgomery wire format, kinds, selectors,
     copies, constants, lookups), a handful of explicit field elements, and the mock-prover verdicts.

The digests pin the oracle (and, through the parity tests, the CUDA path and the C++ host mirror) against
silent drift between rounds.  They are NOT outputs of the Rust reference, which cannot be built here
(no cargo/rustc, un-vendored crates): parity with the reference itself stays "unpinned" (DESIGN.md section 5).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import runpy
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REFERENCE = "/root/reference"
SEED = 20261018
GAMMA = 0x1234567890ABCDEF1234567890ABCDEF0123456789ABCDEF % po.R_MOD


def run_input_creator(n: int) -> None:
    tmp = tempfile.mkdtemp()
    cwd, argv = os.getcwd(), sys.argv
    try:
        os.chdir(tmp)
        sys.argv = ["input-creator.py", str(n)]
        np.random.seed(SEED)
        runpy.run_path(os.path.join(REFERENCE, "input-creator.py"), run_name="__main__")
        shutil.copy(os.path.join(tmp, "data", "matrix.in"), os.path.join(HERE, f"matrix_{n}x{n}.in"))
        shutil.copy(os.path.join(tmp, "data", "matrix-wrong.in"), os.path.join(HERE, f"matrix-wrong_{n}x{n}.in"))
    finally:
        os.chdir(cwd)
        sys.argv = argv
        shutil.rmtree(tmp)


def context_digest(ctx: po.Context) -> dict:
    h = hashlib.sha256()
    h.update(po.pack_mont(ctx.advice).tobytes())
    adv = h.hexdigest()
    meta = hashlib.sha256()
    meta.update("".join(ctx.kind).encode())
    meta.update(bytes(int(s) for s in ctx.selector))
    meta.update(json.dumps([[list(a), list(b)] for a, b in ctx.copies]).encode())
    meta.update(json.dumps([[i, hex(v)] for i, v in ctx.constants]).encode())
    meta.update(json.dumps(list(ctx.lookups)).encode())
    return {"cells": len(ctx.advice), "witness_cells": ctx.kind.count("W"), "gates": int(sum(ctx.selector)),
            "copies": len(ctx.copies), "constants": len(ctx.constants), "lookups": len(ctx.lookups),
            "advice_sha256": adv, "layout_sha256": meta.hexdigest()}


def svd_circuit(inputs, P, lb, err_size):
    """do_zk_svd's circuit (examples/svd_example.rs:98-201 -> src/svd/mod.rs:32-144)."""
    fp = po.FixedPointChip(P, lb)
    ctx = po.Context(0)
    m, u, v = (po.ZkMatrix.new(ctx, fp, inputs[k]) for k in ("m", "u", "v"))
    d = po.ZkVector.new(ctx, fp, inputs["d"])
    err_svd, err_u = po.err_calc(P, err_size, 100.0, 1e-10, 1e-10)
    out = po.check_svd_phase0(ctx, fp, m, u, v, d, err_svd, err_u, 30)
    ctx1 = po.Context(1)
    gamma = ctx1.load_witness(GAMMA)
    po.check_svd_phase1(ctx1, fp, m, u, v, *out, gamma)
    return ctx, ctx1


def zkmatrix_circuit(a, b, P, lb):
    """README.md:34-47: honest_prover_mat_mul -> rescale_matrix -> (phase 1) verify_mul."""
    fp = po.FixedPointChip(P, lb)
    ctx, ctx1 = po.Context(0), po.Context(1)
    za, zb = po.ZkMatrix.new(ctx, fp, a), po.ZkMatrix.new(ctx, fp, b)
    c_s = po.honest_prover_mat_mul(ctx, za.matrix, zb.matrix)
    c = po.ZkMatrix.rescale_matrix(ctx, fp, c_s)
    ZK = po.ZkMatrix
    ZK.verify_mul(ctx1, fp, za, zb, c_s, ctx1.load_witness(GAMMA))
    return ctx, ctx1, c_s, c


def zkvector_circuit(lb):
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
    return ctx, res


def build_golden() -> dict:
    good = json.load(open(os.path.join(HERE, "matrix_8x8.in")))
    wrong = json.load(open(os.path.join(HERE, "matrix-wrong_8x8.in")))
    g = {"seed": SEED, "gamma": hex(GAMMA), "note": "oracle-model digests; not outputs of the Rust reference"}
    # configs[0]: the SVD circuit on matrix / matrix-wrong
    for name, inp in (("svd_matrix", good), ("svd_matrix_wrong", wrong)):
        ctx, ctx1 = svd_circuit(inp, 42, 19, 8)
        errs = po.mock_prove([ctx, ctx1], 19)
        g[name] = {"P": 42, "lookup_bits": 19, "phase0": context_digest(ctx), "phase1": context_digest(ctx1),
                   "mock_prover_failures": len(errs), "first_failures": errs[:4]}
    # test_zkmatrix: m * v^T of the same fixture
    vt = np.array(good["v"]).T.tolist()
    ctx, ctx1, c_s, c = zkmatrix_circuit(good["m"], vt, 42, 19)
    g["zkmatrix_m_vt"] = {"P": 42, "lookup_bits": 19, "phase0": context_digest(ctx), "phase1": context_digest(ctx1),
                          "mock_prover_failures": len(po.mock_prove([ctx, ctx1], 19)),
                          "c_s_00_mont": hex(po.to_mont(c_s[0][0].value)), "c_s_77_mont": hex(po.to_mont(c_s[7][7].value)),
                          "c_00_mont": hex(po.to_mont(c.matrix[0][0].value)),
                          "gamma_pow_7_mont": hex(po.to_mont(pow(GAMMA, 7, po.R_MOD)))}
    ctx, res = zkvector_circuit(19)
    g["test_zkvector"] = {"P": 32, "lookup_bits": 19, "ctx": context_digest(ctx),
                          "mock_prover_failures": len(po.mock_prove([ctx], 19)),
                          "results_mont": [hex(po.to_mont(r.value)) for r in res]}
    return g


if __name__ == "__main__":
    if os.path.isdir(REFERENCE):
        run_input_creator(8)
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(build_golden(), fh, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))
