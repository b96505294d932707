"""Builds libcompute_cuda.so in-tree with nvcc for sm_100a (no torch extension machinery: the product is a plain
C-ABI shared library).  `python -m compute.scala_b200.build` or `__graft_entry__.build()`."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcompute_cuda.so")
OBJ = os.path.join(HERE, "build")

import sysconfig

HOTCALLS = os.path.join(HERE, "_hotcalls" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
SOURCES = ["ir.cpp", "codegen.cpp", "driver.cpp", "runtime.cpp", "tensor.cpp", "kernels_basic.cu", "gemm_3xtf32.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-I/usr/include"]


def _embed_templates() -> None:
    """jit_templates.cuh -> jit_templates_embed.h (the string NVRTC prepends to every generated kernel)."""
    src = open(os.path.join(CSRC, "jit_templates.cuh")).read()
    out = os.path.join(CSRC, "jit_templates_embed.h")
    assert ')JIT"' not in src
    text = '// generated from jit_templates.cuh by build.py — do not edit\n#pragma once\nstatic const char kJitTemplates[] = R"JIT(\n' + src + ')JIT";\n'
    if not os.path.exists(out) or open(out).read() != text:
        open(out, "w").write(text)


def _stamp() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        h.update(name.encode())
        h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "compute_cuda.h"), "rb").read())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def _build_hotcalls() -> None:
    """csrc/py_hotcalls.c -> _hotcalls.<abi>.so: CPython binding of the two per-step calls (cuda.py uses ctypes for everything else, and
    for these two as well if the interpreter's headers are missing)."""
    inc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(inc, "Python.h")):
        sys.stderr.write("Python.h not found: cuda.py binds its hot calls through ctypes\n")
        return
    cmd = ["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-Wall", "-I", inc, os.path.join(CSRC, "py_hotcalls.c"), "-o", HOTCALLS,
           "-L", HERE, "-lcompute_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("building the _hotcalls extension failed")


def build(force: bool = False, verbose: bool = False) -> str:
    _embed_templates()
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        if not os.path.exists(HOTCALLS):
            _build_hotcalls()
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(OBJ, s + ".o")
        cmd = [NVCC, "-c", os.path.join(CSRC, s), "-o", o] + COMMON + ARCH
        if s.endswith(".cu"):
            cmd += ["-Xptxas", "-v"] if verbose else []
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"---- {s} ----\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(f"---- {s} ----\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed")
    link = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "static", "-lnvrtc", "-ldl", "-lpthread",
                                                        "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    _build_hotcalls()
    open(stamp_file, "w").write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
