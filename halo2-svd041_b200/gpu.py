"""Thin object wrapper over the C ABI (one Handle == one h2svd_ctx == one GPU + stream).

Two families of methods, mirroring the header:
  * host methods take/return numpy uint64[..., 4] arrays (bn256::Fr Montgomery limbs) and go
    through the host-pointer entry points (H2D + kernels + D2H inside the call);
  * `*_dev` methods take torch int64[..., 4] CUDA tensors (same bytes) and only enqueue kernels
    on the handle's stream.
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as ct
from typing import Optional

import numpy as np

from . import _ffi


def _np_fr(*shape) -> np.ndarray:
    return np.empty(shape + (4,), dtype=np.uint64)


class PinnedBuffer:
    """Page-locked host memory from h2svd_host_alloc, viewed as a numpy array (freed with the object)."""

    def __init__(self, shape, dtype=np.uint64) -> None:
        self._lib = _ffi.load()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = ct.c_void_p()
        _ffi.check(self._lib.h2svd_host_alloc(nbytes, ct.byref(p)))
        self._ptr = p
        buf = (ct.c_ubyte * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self) -> None:  # pragma: no cover
        try:
            if self._ptr:
                self.array = None
                self._lib.h2svd_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def _np_ptr(a: np.ndarray) -> ct.c_void_p:
    if a.dtype != np.uint64 and a.dtype != np.float64:
        raise TypeError(f"expected uint64/float64 array, got {a.dtype}")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return ct.c_void_p(a.ctypes.data)


def _fr_shape(a, ndim: int):
    if a.shape[-1] != 4 or a.ndim != ndim + 1:
        raise ValueError(f"expected a [{', '.join('?' * ndim)}, 4] limb array, got {tuple(a.shape)}")
    return a.shape[:-1]


class Handle:
    """h2svd_ctx wrapper.  Not thread-safe (mirrors the reference's `&mut Context`)."""

    def __init__(self, device: int = -1, stream: Optional[int] = None) -> None:
        self._lib = _ffi.load()
        h = ct.c_void_p()
        _ffi.check(self._lib.h2svd_create(ct.byref(h), device, ct.c_void_p(stream or 0)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.h2svd_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self) -> "Handle":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ---- properties ----
    @property
    def device(self) -> int:
        return self._lib.h2svd_device(self._h)

    @property
    def sm_count(self) -> int:
        return self._lib.h2svd_sm_count(self._h)

    @property
    def stream(self) -> int:
        return self._lib.h2svd_stream(self._h) or 0

    @property
    def launch_count(self) -> int:
        return int(self._lib.h2svd_launch_count(self._h))

    def sync(self) -> None:
        _ffi.check(self._lib.h2svd_sync(self._h))

    # ---- triage / tuning switches (per handle; not part of the public header) ----
    TUNE_KEYS = ("matmul_tc", "matmul_small", "matmul_small_width", "matmul_cluster", "matmul_karatsuba", "matmul_streamk", "matmul_variant", "fuse_rescale",
                 "rescale_generic", "matvec_warp_kernel", "matvec_seg", "matvec_segs", "matvec_x2", "rescale_ch", "rescale_store", "rescale_fast_sums", "rescale_ctas", "matvec_coreside", "step_schedule")

    def tune(self, key: str, value: int) -> None:
        """matmul_tc / matmul_small: -1 auto, 0 never, 1 always; matmul_karatsuba: -1 auto, 0 schoolbook, 1..3 variants;
        matmul_streamk: -1 auto, 0 never, 1 always; fuse_rescale, rescale_generic, matvec_warp_kernel: 0/1."""
        _ffi.check(self._lib.h2svd_debug_tune(self._h, key.encode(), int(value)))

    def matmul_timeline(self, enable: bool = True) -> np.ndarray:
        """Triage: returns the [16, 8] %globaltimer stamps CTA 0 of the last tensor-core mat-mul left (rounds x slots:
        0 MMA start, 1 MMAs issued, 2 epilogue sees the accumulators, 3 TMEM read + carry done, 4 accumulators zeroed and
        released, 5 field arithmetic + stores done, 6/7 TMA producer first/last load of the tile) and re-arms (or disarms)."""
        out = np.zeros((16, 8), dtype=np.uint64)
        _ffi.check(self._lib.h2svd_debug_matmul_timeline(self._h, int(enable), _np_ptr(out)))
        return out

    def last_matmul_engine(self) -> str:
        """Which engine the last fr_matmul launch of THIS handle used (bench.py reports the matching roofline)."""
        return {0: "schoolbook", 1: "karatsuba", 2: "tensor", 3: "tensor-small"}.get(
            self._lib.h2svd_debug_last_matmul_engine(self._h), "none")

    # ---- host-pointer entry points (numpy) ----
    def fr_matmul(self, a: np.ndarray, b: np.ndarray, b_transposed: bool = False,
                  out: Optional[np.ndarray] = None) -> np.ndarray:
        n, k = _fr_shape(a, 2)
        if b_transposed:
            m, k2 = _fr_shape(b, 2)
        else:
            k2, m = _fr_shape(b, 2)
        if k != k2:
            raise ValueError("fr_matmul: inner dimensions differ")  # reference :515
        c = _np_fr(n, m) if out is None else out
        _ffi.check(self._lib.h2svd_fr_matmul(self._h, _np_ptr(a), _np_ptr(b), _np_ptr(c), n, k, m,
                                            int(b_transposed)))
        return c

    def freivalds_witness(self, a: np.ndarray, b: np.ndarray, c_s: np.ndarray,
                          gamma: np.ndarray) -> dict:
        n, k = _fr_shape(a, 2)
        k2, m = _fr_shape(b, 2)
        n2, m2 = _fr_shape(c_s, 2)
        if k != k2 or n != n2 or m != m2:  # reference :307-309
            raise ValueError("freivalds_witness: shape mismatch")
        out = dict(powers=_np_fr(m), prefix_cv=_np_fr(n, m), prefix_bv=_np_fr(k, m),
                   prefix_abv=_np_fr(n, k), diff=_np_fr(n), is_zero=_np_fr(n), inv=_np_fr(n))
        g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
        _ffi.check(self._lib.h2svd_freivalds_witness(
            self._h, _np_ptr(a), _np_ptr(b), _np_ptr(c_s), _np_ptr(g), n, k, m,
            _np_ptr(out["powers"]), _np_ptr(out["prefix_cv"]), _np_ptr(out["prefix_bv"]),
            _np_ptr(out["prefix_abv"]), _np_ptr(out["diff"]), _np_ptr(out["is_zero"]),
            _np_ptr(out["inv"])))
        return out

    def rescale_witness_count(self, precision_bits: int, lookup_bits: int, shift_bits: int = -1,
                              a_num_bits: int = -1) -> int:
        w = self._lib.h2svd_rescale_witness_count(precision_bits, lookup_bits, shift_bits, a_num_bits)
        if w < 0:
            _ffi.check(w)
        return w

    def rescale_witness(self, c_s: np.ndarray, precision_bits: int, lookup_bits: int,
                        shift_bits: int = -1, a_num_bits: int = -1, out_q: Optional[np.ndarray] = None,
                        out_wit: Optional[np.ndarray] = None):
        flat = c_s.reshape(-1, 4)
        count = flat.shape[0]
        W = self.rescale_witness_count(precision_bits, lookup_bits, shift_bits, a_num_bits)
        q = _np_fr(count) if out_q is None else out_q
        wit = _np_fr(count, W) if out_wit is None else out_wit
        _ffi.check(self._lib.h2svd_rescale_witness(self._h, _np_ptr(flat), count, precision_bits,
                                                  lookup_bits, shift_bits, a_num_bits, _np_ptr(q),
                                                  _np_ptr(wit)))
        return q.reshape(c_s.shape), wit

    @staticmethod
    def _bnd_limbs(bnd: int) -> np.ndarray:
        return np.array([(bnd >> (64 * i)) & ((1 << 64) - 1) for i in range(4)], dtype=np.uint64)

    def abs_less_than_witness(self, x: np.ndarray, bnd: int, lookup_bits: int,
                              y: Optional[np.ndarray] = None) -> np.ndarray:
        """check_abs_less_than(x [- y], bnd) witnesses (reference src/matrix/mod.rs:425-459)."""
        xf = np.ascontiguousarray(x).reshape(-1, 4)
        yf = np.ascontiguousarray(y).reshape(-1, 4) if y is not None else None
        if yf is not None and yf.shape != xf.shape:
            raise ValueError("abs_less_than_witness: shape mismatch")   # reference :448-449
        b = self._bnd_limbs(bnd)
        W = self._lib.h2svd_abs_less_than_witness_count(_np_ptr(b), lookup_bits, int(y is not None))
        if W < 0:
            _ffi.check(W)
        out = _np_fr(xf.shape[0], W)
        _ffi.check(self._lib.h2svd_abs_less_than_witness(self._h, _np_ptr(xf), _np_ptr(yf) if yf is not None else
                                                        ct.c_void_p(0), xf.shape[0], _np_ptr(b), lookup_bits, _np_ptr(out)))
        return out

    def range_check_witness(self, x: np.ndarray, range_bits: int, lookup_bits: int) -> np.ndarray:
        xf = np.ascontiguousarray(x).reshape(-1, 4)
        W = self._lib.h2svd_range_check_witness_count(range_bits, lookup_bits)
        if W < 0:
            _ffi.check(W)
        out = _np_fr(xf.shape[0], W)
        if W:
            _ffi.check(self._lib.h2svd_range_check_witness(self._h, _np_ptr(xf), xf.shape[0], range_bits, lookup_bits,
                                                          _np_ptr(out)))
        return out

    def mat_times_diag(self, a: np.ndarray, v: np.ndarray) -> np.ndarray:
        rows, lda = _fr_shape(a, 2)
        (cols_v,) = _fr_shape(v, 1)
        out = _np_fr(rows, cols_v)
        _ffi.check(self._lib.h2svd_mat_times_diag(self._h, _np_ptr(a), _np_ptr(v), rows, lda, cols_v, _np_ptr(out)))
        return out

    def zkmatrix_mul_witness(self, a: np.ndarray, b: np.ndarray, gamma: np.ndarray, precision_bits: int,
                             lookup_bits: int, shift_bits: int = -1, a_num_bits: int = -1,
                             bv_rows: Optional[tuple] = None, out: Optional[dict] = None) -> dict:
        """honest_prover_mat_mul -> rescale_matrix -> verify_mul for the rows of `a` in ONE slab-pipelined call
        (h2svd_zkmatrix_mul_witness).  `out` may hold preallocated (e.g. pinned) arrays under the result keys."""
        rows, k = _fr_shape(a, 2)
        k2, m = _fr_shape(b, 2)
        if k != k2:
            raise ValueError("zkmatrix_mul_witness: inner dimensions differ")
        W = self.rescale_witness_count(precision_bits, lookup_bits, shift_bits, a_num_bits)
        r0, r1 = bv_rows if bv_rows is not None else (0, k)
        shapes = dict(c_s=(rows, m), q=(rows, m), wit=(rows * m, W), powers=(m,), prefix_cv=(rows, m),
                      prefix_bv=(r1 - r0, m), prefix_abv=(rows, k), diff=(rows,), is_zero=(rows,), inv=(rows,))
        res = {}
        for key, shp in shapes.items():
            arr = out[key] if out is not None and key in out else _np_fr(*shp)
            if tuple(arr.shape) != shp + (4,):
                raise ValueError(f"zkmatrix_mul_witness: out[{key!r}] has shape {arr.shape}, expected {shp + (4,)}")
            res[key] = arr
        g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
        _ffi.check(self._lib.h2svd_zkmatrix_mul_witness(
            self._h, _np_ptr(a), _np_ptr(b), _np_ptr(g), rows, k, m, precision_bits, lookup_bits, shift_bits,
            a_num_bits, r0, r1, *[_np_ptr(res[key]) for key in ("c_s", "q", "wit", "powers", "prefix_cv", "prefix_bv",
                                                                "prefix_abv", "diff", "is_zero", "inv")]))
        return res

    def zkvec_inner_prefix(self, x: np.ndarray, self_: np.ndarray) -> np.ndarray:
        batch, ln = _fr_shape(x, 2)
        if _fr_shape(self_, 2) != (batch, ln):
            raise ValueError("zkvec_inner_prefix: shape mismatch")  # reference :86
        out = _np_fr(batch, ln)
        _ffi.check(self._lib.h2svd_zkvec_inner_prefix(self._h, _np_ptr(x), _np_ptr(self_), batch, ln,
                                                     _np_ptr(out)))
        return out

    def zkvec_sub(self, self_: np.ndarray, x: np.ndarray) -> np.ndarray:
        if self_.shape != x.shape:
            raise ValueError("zkvec_sub: shape mismatch")  # reference :142
        out = np.empty_like(self_)
        _ffi.check(self._lib.h2svd_zkvec_sub(self._h, _np_ptr(self_), _np_ptr(x), self_.size // 4,
                                            _np_ptr(out)))
        return out

    def isqrt_fixed(self, a: np.ndarray, precision_bits: int) -> np.ndarray:
        out = np.empty_like(a)
        _ffi.check(self._lib.h2svd_isqrt_fixed(self._h, _np_ptr(a), a.size // 4, precision_bits,
                                              _np_ptr(out)))
        return out

    def quantize(self, x: np.ndarray, precision_bits: int) -> np.ndarray:
        xs = np.ascontiguousarray(x, dtype=np.float64)
        out = _np_fr(*xs.shape)
        _ffi.check(self._lib.h2svd_quantize(self._h, _np_ptr(xs), xs.size, precision_bits, _np_ptr(out)))
        return out

    def microbench_imad(self, kind: int, iters: int = 2000) -> float:
        v = ct.c_double()
        _ffi.check(self._lib.h2svd_microbench_imad(self._h, kind, iters, ct.byref(v)))
        return v.value

    def microbench_hbm(self, kind: int, nbytes: int = 2 << 30) -> float:
        """GB/s of a copy (0), streaming writes (1), bulk writes from shared memory (2) or reads (3) over nbytes."""
        v = ct.c_double()
        _ffi.check(self._lib.h2svd_microbench_hbm(self._h, kind, nbytes, ct.byref(v)))
        return v.value

    def microbench_tensor_i8(self, kind: int = 0, min_seconds: float = 0.0) -> float:
        v = ct.c_double()
        _ffi.check(self._lib.h2svd_microbench_tensor_i8(self._h, kind, float(min_seconds), ct.byref(v)))
        return v.value

    # ---- device-pointer entry points (torch CUDA tensors, int64[..., 4]) ----
    @staticmethod
    def _tp(t) -> ct.c_void_p:
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("expected a contiguous CUDA tensor")
        return ct.c_void_p(t.data_ptr())

    def fr_matmul_dev(self, a, b, c, b_transposed: bool = False) -> None:
        n, k = a.shape[0], a.shape[1]
        m = b.shape[0] if b_transposed else b.shape[1]
        _ffi.check(self._lib.h2svd_fr_matmul_dev(self._h, self._tp(a), self._tp(b), self._tp(c), n, k, m,
                                                int(b_transposed)))

    def fr_matmul_naive_dev(self, a, b, c) -> None:
        n, k = a.shape[0], a.shape[1]
        m = b.shape[1]
        _ffi.check(self._lib.h2svd_debug_fr_matmul_naive_dev(self._h, self._tp(a), self._tp(b),
                                                            self._tp(c), n, k, m))

    def freivalds_witness_dev(self, a, b, c_s, gamma, powers, prefix_cv, prefix_bv, prefix_abv, diff,
                              is_zero, inv) -> None:
        n, k = a.shape[0], a.shape[1]
        m = b.shape[1]
        _ffi.check(self._lib.h2svd_freivalds_witness_dev(
            self._h, self._tp(a), self._tp(b), self._tp(c_s), self._tp(gamma), n, k, m,
            self._tp(powers), self._tp(prefix_cv), self._tp(prefix_bv), self._tp(prefix_abv),
            self._tp(diff), self._tp(is_zero), self._tp(inv)))

    def gamma_powers_dev(self, gamma, d: int, out) -> None:
        _ffi.check(self._lib.h2svd_gamma_powers_dev(self._h, self._tp(gamma), d, self._tp(out)))

    def mat_vec_prefix_dev(self, a, v, out, totals=None) -> None:
        rows, ln = a.shape[0], a.shape[1]
        _ffi.check(self._lib.h2svd_mat_vec_prefix_totals_dev(
            self._h, self._tp(a), self._tp(v), rows, ln, self._tp(out),
            self._tp(totals) if totals is not None else ct.c_void_p(0)))

    def mat_vec_prefix_pair_dev(self, a0, out0, totals0, a1, out1, totals1, v) -> None:
        """(a0 . v) and (a1 . v) running sums in one launch; a1 may have zero rows."""
        rows0, ln = a0.shape[0], a0.shape[1]
        rows1 = a1.shape[0]
        null = ct.c_void_p(0)
        _ffi.check(self._lib.h2svd_mat_vec_prefix_pair_dev(
            self._h, self._tp(a0), rows0, self._tp(out0), self._tp(totals0) if totals0 is not None else null,
            self._tp(a1) if rows1 else null, rows1, self._tp(out1) if rows1 else null,
            self._tp(totals1) if (rows1 and totals1 is not None) else null, self._tp(v), ln))

    def gather_dev(self, src, count: int, stride: int, offset: int, out) -> None:
        _ffi.check(self._lib.h2svd_gather_dev(self._h, self._tp(src), count, stride, offset,
                                             self._tp(out)))

    def is_equal_witness_dev(self, x, y, diff, is_zero, inv) -> None:
        _ffi.check(self._lib.h2svd_is_equal_witness_dev(self._h, self._tp(x), self._tp(y), x.shape[0],
                                                       self._tp(diff), self._tp(is_zero), self._tp(inv)))

    def rescale_witness_dev(self, c_s, count: int, precision_bits: int, lookup_bits: int, out_q, out_wit,
                            shift_bits: int = -1, a_num_bits: int = -1) -> None:
        _ffi.check(self._lib.h2svd_rescale_witness_dev(self._h, self._tp(c_s), count, precision_bits,
                                                      lookup_bits, shift_bits, a_num_bits,
                                                      self._tp(out_q), self._tp(out_wit)))

    def fr_matmul_rescale_dev(self, a, b, c_s, precision_bits: int, lookup_bits: int, out_q, out_wit,
                              shift_bits: int = -1, a_num_bits: int = -1) -> None:
        """c_s = a.b and the rescale_matrix witnesses of c_s in one call."""
        n, k, m = a.shape[0], a.shape[1], b.shape[1]
        _ffi.check(self._lib.h2svd_fr_matmul_rescale_dev(self._h, self._tp(a), self._tp(b), n, k, m, precision_bits,
                                                        lookup_bits, shift_bits, a_num_bits, self._tp(c_s),
                                                        self._tp(out_q), self._tp(out_wit)))

    def zkmatrix_mul_witness_dev(self, a, b, gamma, precision_bits: int, lookup_bits: int, c_s, q, wit, powers,
                                 prefix_cv, prefix_bv, prefix_abv, diff, is_zero, inv, bv_rows: Optional[tuple] = None,
                                 shift_bits: int = -1, a_num_bits: int = -1) -> None:
        """honest_prover_mat_mul -> rescale_matrix -> verify_mul for the rows of `a`, device tensors, asynchronous; the
        Freivalds mat-vecs are forked onto the handle's side stream inside the call (h2svd_zkmatrix_mul_witness_dev)."""
        rows, k, m = a.shape[0], a.shape[1], b.shape[1]
        r0, r1 = bv_rows if bv_rows is not None else (0, k)
        _ffi.check(self._lib.h2svd_zkmatrix_mul_witness_dev(
            self._h, self._tp(a), self._tp(b), self._tp(gamma), rows, k, m, precision_bits, lookup_bits, shift_bits,
            a_num_bits, r0, r1, self._tp(c_s), self._tp(q), self._tp(wit), self._tp(powers), self._tp(prefix_cv),
            self._tp(prefix_bv) if r1 > r0 else ct.c_void_p(0), self._tp(prefix_abv), self._tp(diff), self._tp(is_zero),
            self._tp(inv)))

    def mat_vec_totals_dev(self, a, v, totals) -> None:
        rows, ln = a.shape[0], a.shape[1]
        _ffi.check(self._lib.h2svd_mat_vec_totals_dev(self._h, self._tp(a), self._tp(v), rows, ln, self._tp(totals)))

    # ---- CUDA-graph capture of *_dev call sequences ----
    def graph_begin(self) -> None:
        _ffi.check(self._lib.h2svd_graph_begin(self._h))

    def graph_end(self) -> "Graph":
        g = ct.c_void_p()
        _ffi.check(self._lib.h2svd_graph_end(self._h, ct.byref(g)))
        return Graph(self, g)

    def zkvec_inner_prefix_dev(self, x, self_, out) -> None:
        batch, ln = x.shape[0], x.shape[1]
        _ffi.check(self._lib.h2svd_zkvec_inner_prefix_dev(self._h, self._tp(x), self._tp(self_), batch, ln,
                                                         self._tp(out)))

    def zkvec_sub_dev(self, self_, x, out) -> None:
        _ffi.check(self._lib.h2svd_zkvec_sub_dev(self._h, self._tp(self_), self._tp(x),
                                                self_.numel() // 4, self._tp(out)))

    def isqrt_fixed_dev(self, a, precision_bits: int, out) -> None:
        _ffi.check(self._lib.h2svd_isqrt_fixed_dev(self._h, self._tp(a), a.numel() // 4, precision_bits,
                                                  self._tp(out)))

    def quantize_dev(self, x, precision_bits: int, out) -> None:
        _ffi.check(self._lib.h2svd_quantize_dev(self._h, self._tp(x), x.numel(), precision_bits,
                                               self._tp(out)))

    def check_canonical_dev(self, x) -> None:
        _ffi.check(self._lib.h2svd_check_canonical_dev(self._h, self._tp(x), x.numel() // 4))


class Graph:
    """A recorded sequence of *_dev calls (h2svd_graph); launch() replays it on the handle's stream."""

    def __init__(self, handle: Handle, g) -> None:
        self._handle, self._g = handle, g

    def launch(self) -> None:
        _ffi.check(self._handle._lib.h2svd_graph_launch(self._handle._h, self._g))

    def close(self) -> None:
        if self._g:
            self._handle._lib.h2svd_graph_destroy(self._g)
            self._g = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class MultiHandle:
    """h2svd_multi: one handle per GPU in ONE process; rows of A / C partitioned, B replicated, no exchange step."""

    def __init__(self, devices) -> None:
        self._lib = _ffi.load()
        devs = (ct.c_int * len(devices))(*devices)
        h = ct.c_void_p()
        _ffi.check(self._lib.h2svd_multi_create(ct.byref(h), devs, len(devices)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.h2svd_multi_destroy(self._h)
            self._h = None

    def __enter__(self) -> "MultiHandle":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def count(self) -> int:
        return self._lib.h2svd_multi_count(self._h)

    def launch_count(self) -> int:
        return sum(int(self._lib.h2svd_launch_count(self._lib.h2svd_multi_ctx(self._h, i))) for i in range(self.count))

    def zkmatrix_mul_witness(self, a: np.ndarray, b: np.ndarray, gamma: np.ndarray, precision_bits: int,
                             lookup_bits: int, shift_bits: int = -1, a_num_bits: int = -1,
                             out: Optional[dict] = None) -> dict:
        n, k = _fr_shape(a, 2)
        k2, m = _fr_shape(b, 2)
        if k != k2:
            raise ValueError("zkmatrix_mul_witness: inner dimensions differ")
        W = self._lib.h2svd_rescale_witness_count(precision_bits, lookup_bits, shift_bits, a_num_bits)
        if W < 0:
            _ffi.check(W)
        shapes = dict(c_s=(n, m), q=(n, m), wit=(n * m, W), powers=(m,), prefix_cv=(n, m), prefix_bv=(k, m),
                      prefix_abv=(n, k), diff=(n,), is_zero=(n,), inv=(n,))
        res = {}
        for key, shp in shapes.items():
            arr = out[key] if out is not None and key in out else _np_fr(*shp)
            if tuple(arr.shape) != shp + (4,):
                raise ValueError(f"zkmatrix_mul_witness: out[{key!r}] has shape {arr.shape}, expected {shp + (4,)}")
            res[key] = arr
        g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
        _ffi.check(self._lib.h2svd_multi_zkmatrix_mul_witness(
            self._h, _np_ptr(a), _np_ptr(b), _np_ptr(g), n, k, m, precision_bits, lookup_bits, shift_bits, a_num_bits,
            *[_np_ptr(res[key]) for key in ("c_s", "q", "wit", "powers", "prefix_cv", "prefix_bv", "prefix_abv", "diff",
                                            "is_zero", "inv")]))
        return res


# ---- bulk hand-off: cell layouts + value expansion (host functions of the library, no GPU) ----
class _LayoutStruct(ct.Structure):
    _fields_ = [("cells", ct.c_uint32), ("witnesses", ct.c_uint32), ("inputs", ct.c_uint32), ("n_gates", ct.c_uint32),
                ("n_lookups", ct.c_uint32), ("n_copies", ct.c_uint32), ("n_constants", ct.c_uint32),
                ("kind", ct.POINTER(ct.c_uint8)), ("source", ct.POINTER(ct.c_int32)), ("gates", ct.POINTER(ct.c_uint32)),
                ("lookups", ct.POINTER(ct.c_int32)), ("copies", ct.POINTER(ct.c_int32)), ("constants", ct.c_void_p)]


class CellsLayout:
    """h2svd_cells_layout: the static advice-cell structure of ONE unit of an operation (kinds, sources, gate offsets,
    lookup cells, constrain_equal pairs, constants) + expand(): the values of all units in assignment order."""
    WITNESS, CONSTANT, EXISTING = 0, 1, 2

    def __init__(self, ptr) -> None:
        self._lib = _ffi.load()
        self._ptr = ptr
        s = ct.cast(ptr, ct.POINTER(_LayoutStruct)).contents
        self.cells, self.witnesses, self.inputs = s.cells, s.witnesses, s.inputs
        self.kind = [s.kind[i] for i in range(s.cells)]
        self.source = [s.source[i] for i in range(s.cells)]
        self.gates = [s.gates[i] for i in range(s.n_gates)]
        self.lookups = [s.lookups[i] for i in range(s.n_lookups)]
        self.copies = [(s.copies[2 * i], s.copies[2 * i + 1]) for i in range(s.n_copies)]
        buf = (ct.c_uint64 * (4 * s.n_constants)).from_address(s.constants) if s.n_constants else []
        self.constants = np.array(list(buf), dtype=np.uint64).reshape(-1, 4)

    @classmethod
    def _make(cls, fn, *args) -> "CellsLayout":
        p = ct.c_void_p()
        _ffi.check(fn(*args, ct.byref(p)))
        return cls(p)

    @classmethod
    def rescale(cls, precision_bits: int, lookup_bits: int, shift_bits: int = -1, a_num_bits: int = -1) -> "CellsLayout":
        return cls._make(_ffi.load().h2svd_rescale_cells_layout, precision_bits, lookup_bits, shift_bits, a_num_bits)

    @classmethod
    def abs_less_than(cls, bnd: int, lookup_bits: int, with_diff: bool) -> "CellsLayout":
        b = Handle._bnd_limbs(bnd)
        return cls._make(_ffi.load().h2svd_abs_less_than_cells_layout, _np_ptr(b), lookup_bits, int(with_diff))

    @classmethod
    def range_check(cls, range_bits: int, lookup_bits: int) -> "CellsLayout":
        return cls._make(_ffi.load().h2svd_range_check_cells_layout, range_bits, lookup_bits)

    @classmethod
    def is_equal(cls) -> "CellsLayout":
        return cls._make(_ffi.load().h2svd_is_equal_cells_layout)

    def expand(self, inputs: Optional[np.ndarray], wit: Optional[np.ndarray], units: int, threads: int = 0) -> np.ndarray:
        out = _np_fr(units, self.cells)
        null = ct.c_void_p(0)
        _ffi.check(self._lib.h2svd_expand_cells(self._ptr, _np_ptr(inputs) if inputs is not None else null,
                                               _np_ptr(wit) if wit is not None else null, units, _np_ptr(out), threads))
        return out

    def close(self) -> None:
        if self._ptr:
            self._lib.h2svd_cells_layout_destroy(self._ptr)
            self._ptr = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def expand_inner_product_cells(a: np.ndarray, v: np.ndarray, prefix: np.ndarray, per_row_v: bool = False,
                               threads: int = 0) -> np.ndarray:
    rows, ln = _fr_shape(a, 2)
    out = _np_fr(rows, 1 + 3 * ln)
    _ffi.check(_ffi.load().h2svd_expand_inner_product_cells(_np_ptr(a), _np_ptr(v), ln if per_row_v else 0, _np_ptr(prefix),
                                                           rows, ln, _np_ptr(out), threads))
    return out


def expand_gamma_power_cells(gamma: np.ndarray, powers: np.ndarray) -> np.ndarray:
    m = powers.shape[0]
    out = _np_fr(1 + 4 * (m - 1))
    g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
    _ffi.check(_ffi.load().h2svd_expand_gamma_power_cells(_np_ptr(g), _np_ptr(powers), m, _np_ptr(out)))
    return out


def expand_is_equal_cells(x: np.ndarray, y: np.ndarray, diff: np.ndarray, is_zero: np.ndarray, inv: np.ndarray) -> np.ndarray:
    n = x.shape[0]
    out = _np_fr(n, 12)
    _ffi.check(_ffi.load().h2svd_expand_is_equal_cells(_np_ptr(x), _np_ptr(y), _np_ptr(diff), _np_ptr(is_zero), _np_ptr(inv),
                                                      n, _np_ptr(out)))
    return out
