"""TEST INFRASTRUCTURE ONLY — builds oracle/oracle_cpu.c (gcc + OpenMP) into oracle/_build/.
Two variants bracket what FP_CONTRACT ON + -cl-unsafe-math-optimizations allow the reference's OpenCL compiler to do
(OpenCL.scala:1131-1135): `strict` (-ffp-contract=off) and `fma` (-ffp-contract=fast -mfma); they are the CHECKERS.
A third, `native` (-O3 -march=native -fopenmp -ffast-math: the flags BASELINE.md section 4 states plus the host analogue of
-cl-unsafe-math-optimizations, which lets gcc call glibc's vector expf / logf / tanhf as POCL's vectoriser would), is the one bench.py
TIMES as the CPU baseline. -march=native code must not travel between machines, so that variant is built on the machine that runs it
(file name keyed by the CPU's flag set) and never by __graft_entry__.build()."""
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


NATIVE_FLAGS = ["-O3", "-march=native", "-fopenmp", "-ffast-math"]


def _cpu_tag() -> str:
    import hashlib

    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
    except Exception:
        flags = "unknown"
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib_path(variant: str = "strict") -> str:
    if variant == "native":
        return os.path.join(OUT, f"liboracle_cpu_native_{_cpu_tag()}.so")
    return os.path.join(OUT, f"liboracle_cpu_{variant}.so")


def build_native() -> str:
    """the timed CPU baseline: built where it runs (see the module docstring); falls back to the `fma` variant's flags + -ffast-math if
    this gcc rejects -march=native for the host"""
    os.makedirs(OUT, exist_ok=True)
    out = lib_path("native")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    for flags in (NATIVE_FLAGS, ["-O3", "-mavx2", "-mfma", "-fopenmp", "-ffast-math"]):
        r = subprocess.run(["gcc", "-shared", "-fPIC", SRC, "-o", out + ".tmp", "-lm"] + flags, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode == 0:
            os.replace(out + ".tmp", out)
            return out
    raise RuntimeError("building the native oracle variant failed:\n" + r.stdout.decode(errors="replace"))


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
        if variant == "native":
            build_native()
        elif not os.path.exists(lib_path(variant)):
            build()
        L = C.CDLL(lib_path(variant))
        fp, i64 = C.c_void_p, C.c_int64
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
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
