"""gpushare-device-plugin_b200 — B200-native inventory + health path of the GPU-share device plugin.

Layout:
  csrc/        sm_100a HBM-probe kernels + the C ABI (include/gpushare_b200.h) -> libgpushare_b200.so
  _abi.py      ctypes binding of that ABI (the Python twin of the cgo stub in INTEGRATION.md)
  device.py    inventory / arena / probe / cycle / health-event calls
  nvidia/      host-side mirror of the reference's pkg/gpu/nvidia (same names, same behaviour)

Importing this package loads the CUDA library; there is no non-CUDA fallback.
"""
from . import _abi  # noqa: F401  (raises ImportError if libgpushare_b200.so has not been built)

__all__ = ["_abi"]
