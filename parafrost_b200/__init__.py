"""parafrost_b200 -- B200-native SIGmA inprocessing engine behind ParaFROST's solver API.

The product is the C-ABI shared library ``libsigma_b200.so`` (include/sigma.h), hand-written
CUDA for sm_100a.  This package only builds and binds it; there is no CPU or PyTorch fallback:
importing :mod:`parafrost_b200.sigma` raises if the library is missing.
"""
from .build import build, lib_path  # noqa: F401

__all__ = ["build", "lib_path"]
