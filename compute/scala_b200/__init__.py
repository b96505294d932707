"""compute.scala_b200 — B200-native (sm_100a) execution backend for Compute.scala's lazy Tensor API.

    from compute.scala_b200 import cuda
    cuda.init()
    r = cuda.Tensor.tanh(a * b + c).flatArray()

The product is `libcompute_cuda.so` (C ABI in include/compute_cuda.h); `cuda` is a ctypes view of it.
"""
from . import cuda  # noqa: F401
