"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU,
exports every symbol include/h2svd_b200.h declares, and fails loudly (no CPU fallback) when no
GPU is present.  No compute is called here."""
import ctypes as ct
import os
import re

import pytest

from tests.util import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "h2svd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2svd_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    syms = _header_symbols()
    assert len(syms) >= 30
    assert sorted(pkg._ffi.SIGNATURES) == syms


def test_library_exports_every_declared_symbol(pkg):
    build = __import__("importlib").import_module("halo2-svd041_b200.build")
    path = build.build()
    lib = ct.CDLL(path)
    for name in _header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/h2svd_b200.h but not exported"
    assert pkg._ffi.load() is not None
    assert b"sm_100a" in pkg._ffi.load().h2svd_version()


def test_parameter_helpers_need_no_gpu(pkg):
    lib = pkg._ffi.load()
    # W = 4 + 4(n_d + n_r) for n_d, n_r >= 2; a one-limb check_big_less_than_safe has 2 witnesses
    assert lib.h2svd_rescale_witness_count(15, 32, 39, 30) == 8
    assert lib.h2svd_rescale_witness_count(63, 19, -1, -1) == 60
    assert lib.h2svd_rescale_witness_count(32, 19, -1, -1) == 36
    assert lib.h2svd_rescale_witness_count(42, 19, -1, -1) == 44
    assert lib.h2svd_rescale_witness_count(63, 19, 3 * 63, 3 * 63 + 1) == 48
    assert lib.h2svd_rescale_witness_count(64, 19, -1, -1) == pkg._ffi.EINVAL
    assert lib.h2svd_rescale_witness_count(32, 0, -1, -1) == pkg._ffi.EINVAL


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    with pytest.raises(pkg.H2svdError) as ei:
        pkg.Handle()
    assert ei.value.code == pkg._ffi.ENODEV
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under halo2-svd041_b200/ may import, link or load it."""
    pkg_dir = os.path.join(ROOT, "halo2-svd041_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|fr_oracle|libfr_oracle|pyoracle|corac)")
    for dirpath, _, files in os.walk(pkg_dir):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs", ".toml")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} references the oracle"


def test_one_source_list_for_every_build():
    """csrc/SOURCES.txt names every .cu of the library and is what BOTH builds read (build.rs had drifted from
    build.py and missed matmul_tc.cu)."""
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "halo2-svd041_b200", "csrc")
    listed = [ln.strip() for ln in open(os.path.join(csrc, "SOURCES.txt")) if ln.strip() and not ln.startswith("#")]
    assert sorted(listed) == sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))
    b = importlib.import_module("halo2-svd041_b200.build")
    assert b.SOURCES == listed
    rs = open(os.path.join(root, "rust", "h2svd-b200", "build.rs")).read()
    assert "SOURCES.txt" in rs and "matmul.cu" not in rs and "api.cu" not in rs      # no hard-coded list left
