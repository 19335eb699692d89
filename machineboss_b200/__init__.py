"""machineboss_b200: B200-native batched Forward / Backward / Viterbi for Machine Boss transducers.

The product is the CUDA shared library (csrc/, C ABI in include/machineboss_b200.h) and the C++
host mirror of the reference classes (host/).  This Python package only builds the library and
binds its C ABI for tests and benchmarks.
"""
from . import build  # noqa: F401
from . import capi  # noqa: F401

__all__ = ["build", "capi"]
