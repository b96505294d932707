"""ctypes binding of libcompute_cuda.so (include/compute_cuda.h).  Binding only — no arithmetic, no fallback: if the
shared library is missing or cc_init fails (no driver / no sm_100 GPU) an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcompute_cuda.so")
HEADER = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "compute_cuda.h")

u64 = C.c_uint64
i32 = C.c_int32
P = C.POINTER


class ComputeCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[{STATUS_NAMES.get(status, status)}] {message}")
        self.status = status


class IllegalArgumentException(ComputeCudaError, ValueError):
    """java.lang.IllegalArgumentException of the reference API"""


STATUS_NAMES = {
    0: "CC_OK", -1: "CC_ERR_ILLEGAL_ARGUMENT", -2: "CC_ERR_NOT_INITIALIZED", -3: "CC_ERR_NO_DRIVER", -4: "CC_ERR_CUDA",
    -5: "CC_ERR_COMPILE", -6: "CC_ERR_BAD_TREE", -7: "CC_ERR_NCCL", -8: "CC_ERR_UNSUPPORTED", -9: "CC_ERR_OUT_OF_MEMORY",
}


class DeviceInfo(C.Structure):
    _fields_ = [("ordinal", i32), ("sm_count", i32), ("cc_major", i32), ("cc_minor", i32), ("max_smem_per_block", i32),
                ("l2_bytes", i32), ("total_mem", C.c_int64), ("sm_clock_khz", i32), ("mem_clock_khz", i32), ("name", C.c_char * 64)]


class KernelInfo(C.Structure):
    _fields_ = [("kind", i32), ("cache_hit", i32), ("n_args", i32), ("n_launches", i32), ("out_floats", u64),
                ("algorithmic_bytes", u64), ("flops", u64), ("structural_hash", u64)]


class Stats(C.Structure):
    _fields_ = [(n, u64) for n in ("compiles", "cache_hits", "launches", "device_kernels", "h2d_bytes", "d2h_bytes", "alloc_calls",
                                   "pool_hits", "bytes_in_use", "bytes_pooled", "nvrtc_compiles", "disk_cache_hits")]


def declared_symbols() -> list[str]:
    """every function include/compute_cuda.h declares"""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cc|ct)_[a-z0-9_]+)\s*\(", text)) - {"cc_event_callback"})


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run `python -m compute.scala_b200.build` (there is no fallback path)")
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.cc_last_error.restype = C.c_char_p
        _lib.cc_version.restype = C.c_char_p
        for name in declared_symbols():
            fn = getattr(_lib, name)  # AttributeError if the library does not export a declared symbol
            if name not in ("cc_last_error", "cc_version"):
                fn.restype = C.c_int
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().cc_last_error().decode("utf-8", "replace")
        if status == -1:
            raise IllegalArgumentException(status, msg)
        raise ComputeCudaError(status, msg)
