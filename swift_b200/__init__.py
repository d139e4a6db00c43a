"""swift_b200 - B200-native SPH neighbour-interaction path for SWIFT.

The product is the CUDA shared library ``libswiftgpu.so`` (C ABI declared in
``include/swiftgpu.h``). This package holds its sources (``csrc/``), the ctypes
mirror of the ABI (``abi``), a thin host-side wrapper (``engine.SwiftGPU``) and
the synthetic-input harness (``host``).
"""
from . import abi  # noqa: F401
