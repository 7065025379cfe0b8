"""B200-native batched quintic motion-primitive generation + convex-obstacle collision checking.

The product is the CUDA C-ABI library (csrc/ -> librqcd.so, include/rqcd.h) and the C++ facade in
include/rqcd/.  The Python here is harness plumbing for tests/ and bench.py.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
