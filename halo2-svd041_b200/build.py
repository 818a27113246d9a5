"""Builds libh2svd_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

    python halo2-svd041_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("H2SVD_OBJ_DIR") or os.path.join(HERE, "build")
LIB = os.environ.get("H2SVD_LIB_OUT") or os.path.join(HERE, "libh2svd_b200.so")   # (the two overrides: tuning builds)
# one list for every build of the library: this script and rust/h2svd-b200/build.rs both read csrc/SOURCES.txt
with open(os.path.join(CSRC, "SOURCES.txt")) as _fh:
    SOURCES = [ln.strip() for ln in _fh if ln.strip() and not ln.startswith("#")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "h2svd_b200.h"))
    return max(os.path.getmtime(f) for f in files)


# per-file extra flags (tuning hook: H2SVD_EXTRA_FLAGS_matmul="-Xptxas -O1")
def _extra(src: str):
    return os.environ.get("H2SVD_EXTRA_FLAGS_" + src.replace(".cu", ""), "").split()


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    log = obj + ".log"
    cmd = [NVCC, *FLAGS, *_extra(src), "-c", os.path.join(CSRC, src), "-o", obj]
    with open(log, "w") as fh:
        rc = subprocess.call(cmd, stdout=fh, stderr=subprocess.STDOUT)
    if rc != 0:
        sys.stderr.write(open(log).read())
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
