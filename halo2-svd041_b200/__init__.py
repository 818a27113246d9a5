"""halo2-svd041_b200 -- B200-native ZkMatrix / ZkVector witness path (sm_100a CUDA behind a C ABI).

The directory name is not a Python identifier; import it with
    importlib.import_module("halo2-svd041_b200")
Importing the package does NOT load the CUDA library (so CPU-only tooling can inspect it);
`Handle()` / `_ffi.load()` do, and fail loudly if it has not been built.  There is no CPU fallback.
"""
from . import _ffi  # noqa: F401
from ._ffi import H2svdError  # noqa: F401
from .gpu import Handle, PinnedBuffer, last_matmul_engine, set_matmul_karatsuba, set_matmul_streamk, set_matmul_tc, set_fuse_rescale, set_matmul_variant, set_matvec_warp_kernel, set_rescale_generic  # noqa: F401

__all__ = ["Handle", "PinnedBuffer", "H2svdError", "last_matmul_engine", "set_matmul_karatsuba", "set_matmul_streamk", "set_matmul_tc", "set_fuse_rescale", "set_matmul_variant", "set_matvec_warp_kernel", "set_rescale_generic"]
