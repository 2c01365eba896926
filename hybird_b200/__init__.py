"""hybird_b200 -- B200-native engine for the lattice-Boltzmann hot path of gnomeCreative/hybird.

Product code: the CUDA library (csrc/ -> liblbgpu.so, C ABI in include/lbgpu.h), its ctypes
binding (abi.py), the host mirror of the reference's LB class (lb.py), the box-domain
initialiser (lattice_init.py) and the C++ drop-in shim for the reference driver (shim/).
Nothing in this package imports the CPU oracles under oracle/.
"""
from .abi import LbGpuError, load_library  # noqa: F401
from .lb import LB  # noqa: F401

__all__ = ["LB", "LbGpuError", "load_library"]
