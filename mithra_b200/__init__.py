"""mithra_b200: B200-native MITHRA FDTD/PIC time-march.

The product is libmithra_gpu.so (CUDA kernels behind the C ABI of include/mithra_gpu.h) and the C++ host
classes in mithra_b200/host.  This Python package is the ctypes harness used by tests/ and bench.py.
"""
from . import abi  # noqa: F401
