"""cfs_spmv_b200 -- B200-native symmetric SpMV behind the cfs-spmv C++ API.

The product is native: CUDA kernels + a C ABI (csrc/, include/cfs_cuda.h) and a
C++ host library with the reference's class names (host/, include/cfs.hpp).
This Python package only builds them and binds the C ABI for tests and bench.
"""
from . import capi  # noqa: F401
