"""ctypes binding of libh2svd_b200.so (include/h2svd_b200.h).

The library is the product; this file only declares signatures.  Loading fails loudly when the
CUDA library has not been built -- there is no CPU fallback and the oracle is never imported here.
"""
from __future__ import annotations

import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# H2SVD_LIB: tuning only -- load another build of the same library (e.g. one compiled with experiment macros)
LIB_PATH = os.environ.get("H2SVD_LIB") or os.path.join(_HERE, "libh2svd_b200.so")

OK, EINVAL, ECUDA, ENOMEM, ENODEV, ERANGE = 0, -1, -2, -3, -4, -5
_CODES = {EINVAL: "EINVAL", ECUDA: "ECUDA", ENOMEM: "ENOMEM", ENODEV: "ENODEV", ERANGE: "ERANGE"}


class H2svdError(RuntimeError):
    def __init__(self, code: int, message: str) -> None:
        super().__init__(f"h2svd error {_CODES.get(code, code)}: {message}")
        self.code = code


_P, _Z, _I = ct.c_void_p, ct.c_size_t, ct.c_int

# name -> (restype, argtypes); every symbol include/h2svd_b200.h declares
SIGNATURES = {
    "h2svd_last_error": (ct.c_char_p, []),
    "h2svd_version": (ct.c_char_p, []),
    "h2svd_create": (_I, [ct.POINTER(_P), _I, _P]),
    "h2svd_destroy": (None, [_P]),
    "h2svd_sync": (_I, [_P]),
    "h2svd_stream": (_P, [_P]),
    "h2svd_device": (_I, [_P]),
    "h2svd_sm_count": (_I, [_P]),
    "h2svd_launch_count": (ct.c_uint64, [_P]),
    "h2svd_fr_matmul": (_I, [_P, _P, _P, _P, _Z, _Z, _Z, _I]),
    "h2svd_fr_matmul_dev": (_I, [_P, _P, _P, _P, _Z, _Z, _Z, _I]),
    "h2svd_freivalds_witness": (_I, [_P, _P, _P, _P, _P, _Z, _Z, _Z, _P, _P, _P, _P, _P, _P, _P]),
    "h2svd_freivalds_witness_dev": (_I, [_P, _P, _P, _P, _P, _Z, _Z, _Z, _P, _P, _P, _P, _P, _P, _P]),
    "h2svd_gamma_powers_dev": (_I, [_P, _P, _Z, _P]),
    "h2svd_mat_vec_prefix_dev": (_I, [_P, _P, _P, _Z, _Z, _P]),
    "h2svd_mat_vec_prefix_totals_dev": (_I, [_P, _P, _P, _Z, _Z, _P, _P]),
    "h2svd_mat_vec_prefix_pair_dev": (_I, [_P, _P, _Z, _P, _P, _P, _Z, _P, _P, _P, _Z]),
    "h2svd_mat_vec_prefix": (_I, [_P, _P, _P, _Z, _Z, _P]),
    "h2svd_host_fr_from_canonical": (_I, [_P, _P]),
    "h2svd_host_fr_to_canonical": (None, [_P, _P]),
    "h2svd_host_fr_add": (None, [_P, _P, _P]),
    "h2svd_host_fr_sub": (None, [_P, _P, _P]),
    "h2svd_host_fr_mul": (None, [_P, _P, _P]),
    "h2svd_gather_dev": (_I, [_P, _P, _Z, _Z, _Z, _P]),
    "h2svd_is_equal_witness_dev": (_I, [_P, _P, _P, _Z, _P, _P, _P]),
    "h2svd_abs_less_than_witness_count": (_I, [_P, _I, _I]),
    "h2svd_abs_less_than_witness": (_I, [_P, _P, _P, _Z, _P, _I, _P]),
    "h2svd_abs_less_than_witness_dev": (_I, [_P, _P, _P, _Z, _P, _I, _P]),
    "h2svd_range_check_witness_count": (_I, [_I, _I]),
    "h2svd_range_check_witness": (_I, [_P, _P, _Z, _I, _I, _P]),
    "h2svd_range_check_witness_dev": (_I, [_P, _P, _Z, _I, _I, _P]),
    "h2svd_mat_times_diag": (_I, [_P, _P, _P, _Z, _Z, _Z, _P]),
    "h2svd_mat_times_diag_dev": (_I, [_P, _P, _P, _Z, _Z, _Z, _P]),
    "h2svd_zkmatrix_mul_witness": (_I, [_P, _P, _P, _P, _Z, _Z, _Z, _I, _I, _I, _I, _Z, _Z] + [_P] * 10),
    "h2svd_zkmatrix_mul_witness_dev": (_I, [_P, _P, _P, _P, _Z, _Z, _Z, _I, _I, _I, _I, _Z, _Z] + [_P] * 10),
    "h2svd_mat_vec_totals_dev": (_I, [_P, _P, _P, _Z, _Z, _P]),
    "h2svd_graph_begin": (_I, [_P]),
    "h2svd_graph_end": (_I, [_P, ct.POINTER(_P)]),
    "h2svd_graph_launch": (_I, [_P, _P]),
    "h2svd_graph_destroy": (None, [_P]),
    "h2svd_multi_create": (_I, [ct.POINTER(_P), ct.POINTER(_I), _I]),
    "h2svd_multi_destroy": (None, [_P]),
    "h2svd_multi_count": (_I, [_P]),
    "h2svd_multi_ctx": (_P, [_P, _I]),
    "h2svd_multi_zkmatrix_mul_witness": (_I, [_P, _P, _P, _P, _Z, _Z, _Z, _I, _I, _I, _I] + [_P] * 10),
    "h2svd_host_alloc": (_I, [_Z, ct.POINTER(_P)]),
    "h2svd_host_free": (None, [_P]),
    "h2svd_rescale_witness_count": (_I, [_I, _I, _I, _I]),
    "h2svd_rescale_witness": (_I, [_P, _P, _Z, _I, _I, _I, _I, _P, _P]),
    "h2svd_rescale_witness_dev": (_I, [_P, _P, _Z, _I, _I, _I, _I, _P, _P]),
    "h2svd_fr_matmul_rescale_dev": (_I, [_P, _P, _P, _Z, _Z, _Z, _I, _I, _I, _I, _P, _P, _P]),
    "h2svd_zkvec_inner_prefix": (_I, [_P, _P, _P, _Z, _Z, _P]),
    "h2svd_zkvec_inner_prefix_dev": (_I, [_P, _P, _P, _Z, _Z, _P]),
    "h2svd_zkvec_sub": (_I, [_P, _P, _P, _Z, _P]),
    "h2svd_zkvec_sub_dev": (_I, [_P, _P, _P, _Z, _P]),
    "h2svd_isqrt_fixed": (_I, [_P, _P, _Z, _I, _P]),
    "h2svd_isqrt_fixed_dev": (_I, [_P, _P, _Z, _I, _P]),
    "h2svd_quantize": (_I, [_P, _P, _Z, _I, _P]),
    "h2svd_quantize_dev": (_I, [_P, _P, _Z, _I, _P]),
    "h2svd_check_canonical_dev": (_I, [_P, _P, _Z]),
    "h2svd_rescale_cells_layout": (_I, [_I, _I, _I, _I, ct.POINTER(_P)]),
    "h2svd_abs_less_than_cells_layout": (_I, [_P, _I, _I, ct.POINTER(_P)]),
    "h2svd_range_check_cells_layout": (_I, [_I, _I, ct.POINTER(_P)]),
    "h2svd_is_equal_cells_layout": (_I, [ct.POINTER(_P)]),
    "h2svd_cells_layout_destroy": (None, [_P]),
    "h2svd_expand_cells": (_I, [_P, _P, _P, _Z, _P, _I]),
    "h2svd_expand_inner_product_cells": (_I, [_P, _P, _Z, _P, _Z, _Z, _P, _I]),
    "h2svd_expand_gamma_power_cells": (_I, [_P, _P, _Z, _P]),
    "h2svd_expand_is_equal_cells": (_I, [_P, _P, _P, _P, _P, _Z, _P]),
    "h2svd_microbench_imad": (_I, [_P, _I, _I, ct.POINTER(ct.c_double)]),
    "h2svd_microbench_hbm": (_I, [_P, _I, _Z, ct.POINTER(ct.c_double)]),
    "h2svd_microbench_tensor_i8": (_I, [_P, _I, ct.c_double, ct.POINTER(ct.c_double)]),
}
# not part of the public header: triage helpers
DEBUG_SIGNATURES = {
    "h2svd_debug_fr_matmul_naive_dev": (_I, [_P, _P, _P, _P, _Z, _Z, _Z]),
    "h2svd_debug_tune": (_I, [_P, ct.c_char_p, _I]),          # per-handle tuning switches
    "h2svd_debug_last_matmul_engine": (_I, [_P]),
    "h2svd_debug_matmul_timeline": (_I, [_P, _I, _P]),
}

_LIB = None


def load() -> ct.CDLL:
    """Loads the CUDA library; raises if it is missing (no fallback of any kind)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python halo2-svd041_b200/build.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ct.CDLL(LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **DEBUG_SIGNATURES}.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise H2svdError(rc, load().h2svd_last_error().decode(errors="replace"))
