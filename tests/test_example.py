"""examples/svd_example.cpp -- the C++ counterpart of the reference's examples/svd_example.rs -- on the committed
reference-generated fixtures: `matrix` verifies (exit 0), `matrix-wrong` is rejected (exit 1) (README.md:93 of the
reference), and the two smoke drivers (tasks 1, 2) run to a satisfied constraint system."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import ROOT

pytestmark = pytest.mark.gpu
EX = os.path.join(ROOT, "examples")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-C", EX, "-B", "svd_example"], stdout=subprocess.DEVNULL)
    return os.path.join(EX, "svd_example")


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("data")
    gold = os.path.join(ROOT, "tests", "golden")
    shutil.copy(os.path.join(gold, "matrix_8x8.in"), d / "matrix.in")
    shutil.copy(os.path.join(gold, "matrix-wrong_8x8.in"), d / "matrix-wrong.in")
    good, wrong = po.make_svd_inputs(48, 48, 99)          # a larger input-creator.py style case
    for name, inp in (("big", good), ("big-wrong", wrong)):
        with open(d / f"{name}.in", "w") as fh:
            json.dump({k: np.asarray(v).tolist() for k, v in inp.items()}, fh)
    return str(d)


def _run(exe, *args):
    p = subprocess.run([exe, *args], capture_output=True, text=True, timeout=300)
    return p.returncode, p.stdout + p.stderr


def test_matrix_verifies_and_matrix_wrong_fails(exe, data_dir):
    rc, out = _run(exe, "matrix", "--data-dir", data_dir)
    assert rc == 0 and "VERIFIED" in out, out
    rc, out = _run(exe, "matrix-wrong", "--data-dir", data_dir)
    assert rc == 1 and "VIOLATED" in out and "ctx 0" in out and "ctx 1" not in out, out   # rejected in phase 0
    rc, out = _run(exe, "big", "--data-dir", data_dir)
    assert rc == 0, out
    rc, out = _run(exe, "big-wrong", "--data-dir", data_dir)
    assert rc == 1, out


def test_smoke_drivers(exe, data_dir):
    rc, out = _run(exe, "--task", "1")
    assert rc == 0 and "Inner product" in out, out
    rc, out = _run(exe, "--task", "2")
    assert rc == 0, out


def test_usage_and_missing_file(exe, data_dir):
    assert _run(exe)[0] == 2
    rc, out = _run(exe, "nope", "--data-dir", data_dir)
    assert rc == 2 and "Unable to read file" in out
