"""TEST INFRASTRUCTURE ONLY — builds oracle/oracle_cpu.c (gcc + OpenMP) into oracle/_build/.
Two variants bracket what FP_CONTRACT ON + -cl-unsafe-math-optimizations allow the reference's OpenCL compiler to do
(OpenCL.scala:1131-1135): `strict` (-ffp-contract=off) and `fma` (-ffp-contract=fast -mfma)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "oracle_cpu.c")
VARIANTS = {
    "strict": ["-O3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math"],
    "fma": ["-O3", "-fopenmp", "-ffp-contract=fast", "-mfma", "-mavx2", "-fno-fast-math"],
}


def lib_path(variant: str = "strict") -> str:
    return os.path.join(OUT, f"liboracle_cpu_{variant}.so")


def build(force: bool = False) -> None:
    os.makedirs(OUT, exist_ok=True)
    for v, flags in VARIANTS.items():
        out = lib_path(v)
        if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
            continue
        subprocess.run(["gcc", "-shared", "-fPIC", SRC, "-o", out, "-lm"] + flags, check=True)


_libs: dict = {}


def load(variant: str = "strict") -> C.CDLL:
    if variant not in _libs:
        if not os.path.exists(lib_path(variant)):
            build()
        L = C.CDLL(lib_path(variant))
        fp, i64 = C.c_void_p, C.c_int64
        L.oracle_num_threads.restype = C.c_int
        L.oracle_random.argtypes = [fp, i64, C.c_uint32]
        L.oracle_c1.argtypes = [fp, fp, fp, fp, i64]
        L.oracle_c2.argtypes = [fp, fp, fp, fp, i64]
        L.oracle_sum_cpu_order.argtypes = [fp, i64]
        L.oracle_sum_cpu_order.restype = C.c_float
        L.oracle_sum_fp64.argtypes = [fp, i64]
        L.oracle_sum_fp64.restype = C.c_double
        L.oracle_axis_sum_2d.argtypes = [fp, i64, i64, C.c_int, fp]
        L.oracle_matmul_left_fold.argtypes = [fp, fp, fp, i64, i64, i64]
        L.oracle_affine_gather_3d.argtypes = [fp, fp, C.c_int, fp, fp, C.c_float, fp]
        _libs[variant] = L
    return _libs[variant]


if __name__ == "__main__":
    build(force=True)
    print(lib_path("strict"), lib_path("fma"))
