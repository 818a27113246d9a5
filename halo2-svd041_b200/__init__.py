"""halo2-svd041_b200 -- B200-native ZkMatrix / ZkVector witness path (sm_100a CUDA behind a C ABI).

The directory name is not a Python identifier; import it with
    importlib.import_module("halo2-svd041_b200")
Importing the package does NOT load the CUDA library (so CPU-only tooling can inspect it);
`Handle()` / `_ffi.load()` do, and fail loudly if it has not been built.  There is no CPU fallback.
"""
from . import _ffi  # noqa: F401
from ._ffi import H2svdError  # noqa: F401
from .gpu import (CellsLayout, Graph, Handle, MultiHandle, PinnedBuffer, expand_gamma_power_cells,  # noqa: F401
                  expand_inner_product_cells, expand_is_equal_cells)

__all__ = ["Handle", "MultiHandle", "Graph", "PinnedBuffer", "H2svdError", "CellsLayout", "expand_inner_product_cells",
           "expand_gamma_power_cells", "expand_is_equal_cells"]
