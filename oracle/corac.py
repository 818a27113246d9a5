"""ctypes loader for the C oracle (oracle/libfr_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  Arrays are numpy uint64[..., 4]
(halo2curves Montgomery limbs)."""
from __future__ import annotations

import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libfr_oracle.so")
    src = os.path.join(_HERE, "fr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfr_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ct.CDLL(build())
        P, Z, I = ct.c_void_p, ct.c_size_t, ct.c_int
        L = _LIB
        L.orc_field_mat_mul.argtypes = [P, P, P, Z, Z, Z, Z, Z, I]
        L.orc_gamma_powers.argtypes = [P, Z, P]
        L.orc_mat_vec_prefix.argtypes = [P, P, Z, Z, P, I]
        L.orc_rescale_witness_count.argtypes = [I, I, I, I]
        L.orc_rescale_witness.argtypes = [P, Z, I, I, I, I, P, P, P, I]
        L.orc_zkvec_inner_prefix.argtypes = [P, P, Z, Z, P, I]
        L.orc_zkvec_sub.argtypes = [P, P, Z, P]
        L.orc_quantize.argtypes = [P, Z, I, P]
        L.orc_isqrt_fixed.argtypes = [P, Z, I, P]
        L.orc_freivalds_witness.argtypes = [P, P, P, P, Z, Z, Z, P, P, P, P, P, P, P, I]
        L.orc_abs_less_than_witness_count.argtypes = [P, I, I]
        L.orc_abs_less_than_witness.argtypes = [P, P, Z, P, I, P]
        L.orc_range_check_witness_count.argtypes = [I, I]
        L.orc_range_check_witness.argtypes = [P, Z, I, I, P]
        L.orc_mat_times_diag.argtypes = [P, P, Z, Z, Z, P]
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ct.c_void_p)


def _fr(*shape) -> np.ndarray:
    return np.zeros(shape + (4,), dtype=np.uint64)


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed: {rc}")


def field_mat_mul(a: np.ndarray, b: np.ndarray, row0: int = 0, row1: int | None = None,
                  threads: int = 1) -> np.ndarray:
    n, k = a.shape[:2]
    k2, m = b.shape[:2]
    assert k == k2
    row1 = n if row1 is None else row1
    c = _fr(n, m)
    _chk(lib().orc_field_mat_mul(_p(a), _p(b), _p(c), n, k, m, row0, row1, threads), "mat_mul")
    return c


def gamma_powers(gamma: np.ndarray, d: int) -> np.ndarray:
    out = _fr(d)
    _chk(lib().orc_gamma_powers(_p(gamma), d, _p(out)), "powers")
    return out


def mat_vec_prefix(a: np.ndarray, v: np.ndarray, threads: int = 1) -> np.ndarray:
    rows, ln = a.shape[:2]
    out = _fr(rows, ln)
    _chk(lib().orc_mat_vec_prefix(_p(a), _p(v), rows, ln, _p(out), threads), "mat_vec")
    return out


def rescale_witness_count(P: int, lb: int, S: int = -1, A: int = -1) -> int:
    return lib().orc_rescale_witness_count(P, lb, S, A)


def rescale_witness(cs: np.ndarray, P: int, lb: int, S: int = -1, A: int = -1, threads: int = 1):
    flat = np.ascontiguousarray(cs).reshape(-1, 4)
    count = flat.shape[0]
    W = rescale_witness_count(P, lb, S, A)
    q, rem, wit = _fr(count), _fr(count), _fr(count, W)
    _chk(lib().orc_rescale_witness(_p(flat), count, P, lb, S, A, _p(q), _p(rem), _p(wit), threads),
         "rescale")
    return q, rem, wit


def zkvec_inner_prefix(x: np.ndarray, self_: np.ndarray, threads: int = 1) -> np.ndarray:
    batch, ln = x.shape[:2]
    out = _fr(batch, ln)
    _chk(lib().orc_zkvec_inner_prefix(_p(x), _p(self_), batch, ln, _p(out), threads), "inner")
    return out


def zkvec_sub(self_: np.ndarray, x: np.ndarray) -> np.ndarray:
    flat = self_.reshape(-1, 4)
    out = _fr(flat.shape[0])
    _chk(lib().orc_zkvec_sub(_p(flat), _p(x.reshape(-1, 4)), flat.shape[0], _p(out)), "sub")
    return out.reshape(self_.shape)


def quantize(x: np.ndarray, P: int) -> np.ndarray:
    xs = np.ascontiguousarray(x, dtype=np.float64)
    out = _fr(*xs.shape)
    _chk(lib().orc_quantize(_p(xs), xs.size, P, _p(out)), "quantize")
    return out


def isqrt_fixed(a: np.ndarray, P: int) -> np.ndarray:
    flat = np.ascontiguousarray(a).reshape(-1, 4)
    out = _fr(flat.shape[0])
    _chk(lib().orc_isqrt_fixed(_p(flat), flat.shape[0], P, _p(out)), "isqrt")
    return out


def freivalds_witness(a: np.ndarray, b: np.ndarray, cs: np.ndarray, gamma: np.ndarray,
                      threads: int = 1) -> dict:
    n, k = a.shape[:2]
    _, m = b.shape[:2]
    out = dict(powers=_fr(m), prefix_cv=_fr(n, m), prefix_bv=_fr(k, m), prefix_abv=_fr(n, k),
               diff=_fr(n), is_zero=_fr(n), inv=_fr(n))
    _chk(lib().orc_freivalds_witness(_p(a), _p(b), _p(cs), _p(gamma), n, k, m, _p(out["powers"]),
                                     _p(out["prefix_cv"]), _p(out["prefix_bv"]),
                                     _p(out["prefix_abv"]), _p(out["diff"]), _p(out["is_zero"]),
                                     _p(out["inv"]), threads), "freivalds")
    return out


def _bnd(bnd: int) -> np.ndarray:
    return np.array([(bnd >> (64 * i)) & ((1 << 64) - 1) for i in range(4)], dtype=np.uint64)


def abs_less_than_witness(x: np.ndarray, bnd: int, lb: int, y: np.ndarray | None = None) -> np.ndarray:
    xf = np.ascontiguousarray(x).reshape(-1, 4)
    b = _bnd(bnd)
    W = lib().orc_abs_less_than_witness_count(_p(b), lb, int(y is not None))
    if W < 0:
        raise ValueError("abs_less_than: parameters out of range")
    out = _fr(xf.shape[0], W)
    yf = np.ascontiguousarray(y).reshape(-1, 4) if y is not None else None
    _chk(lib().orc_abs_less_than_witness(_p(xf), _p(yf) if yf is not None else None, xf.shape[0], _p(b), lb, _p(out)),
         "abs_less_than_witness")
    return out


def range_check_witness(x: np.ndarray, range_bits: int, lb: int) -> np.ndarray:
    xf = np.ascontiguousarray(x).reshape(-1, 4)
    W = lib().orc_range_check_witness_count(range_bits, lb)
    if W < 0:
        raise ValueError("range_check: parameters out of range")
    out = _fr(xf.shape[0], W)
    if W:
        _chk(lib().orc_range_check_witness(_p(xf), xf.shape[0], range_bits, lb, _p(out)), "range_check_witness")
    return out


def mat_times_diag(a: np.ndarray, v: np.ndarray) -> np.ndarray:
    rows, lda = a.shape[0], a.shape[1]
    out = _fr(rows, v.shape[0])
    _chk(lib().orc_mat_times_diag(_p(a), _p(v), rows, lda, v.shape[0], _p(out)), "mat_times_diag")
    return out
