"""B200-native block texture compressor behind FasTC's Core API.

The product is libfastc_gpu.so (hand-written sm_100a CUDA kernels + a C ABI,
see include/fastc_gpu.h) and the C++ host layer in fastc_b200/core/ that mirrors
FasTC's `CompressImageData` / `CompressionJob` / `SCompressionSettings`.  This
Python package is a thin ctypes mirror of the same interface, used by the
parity tests and bench.py.  There is no CPU fallback anywhere in this package:
if the CUDA library is missing or no GPU is visible, calls raise.
"""
from .api import (  # noqa: F401
    ECompressionFormat,
    SCompressionSettings,
    CompressImageData,
    CompressedImage,
    FastcGpuError,
    GpuLibrary,
    lib,
)
