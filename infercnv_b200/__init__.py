"""infercnv_b200 - B200 (sm_100a) engine for inferCNV's smoothing + HMM hot path.

The product is `libinfercnv_b200.so` (C ABI in include/infercnv_b200.h, CUDA in csrc/).
This package is the host-side mirror of the reference's R interface for that path:

* `infercnv_b200.api`      NumPy wrappers of the host-pointer ABI (what the R shim binds)
* `infercnv_b200.ops`      same names / arguments / error behaviour as the R functions they replace
* `infercnv_b200.device`   torch-tensor wrappers of the device-pointer ABI, sharding over GPUs
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
