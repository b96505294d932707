"""TEST INFRASTRUCTURE ONLY (see README.md): compile a generated kernel for the host and run it on small grids."""
import ctypes as C
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = os.path.join(tempfile.gettempdir(), "compute_cuda_kernel_emulator")
ARG_OUT, ARG_SCRATCH0, ARG_PARTIALS, ARG_COUNTER, ARG_COL_COUNTERS, ARG_PEER_MB, ARG_PEER_EPOCH = -1, -2, -100, -101, -102, -103, -104
COL_COUNTERS = np.zeros(4096, np.uint32)  # persistent and self-resetting, like the runtime's
FOLD_PARTIALS = 148 * 4 * 2


class LaunchInfo(C.Structure):
    _fields_ = [("entry", C.c_char * 64), ("grid", C.c_uint32 * 3), ("block", C.c_uint32 * 3), ("smem", C.c_uint32), ("n_args", C.c_int32),
                ("args", C.c_int32 * 32), ("n_scratch", C.c_int32), ("scratch_floats", C.c_uint64 * 8)]


def launches_of(cuda, kernel):
    L = cuda._L()
    L.cc_kernel_launch_info.argtypes = [C.c_uint64, C.c_int, C.POINTER(LaunchInfo)]
    out = []
    info = kernel.info
    for i in range(info.n_launches):
        li = LaunchInfo()
        st = L.cc_kernel_launch_info(kernel.handle, i, C.byref(li))
        if st != 0 and info.kind == 2:
            break  # a contraction's n_launches also counts the precompiled tcgen05 kernels, which have no generated entry
        cuda.check(st)
        out.append(li)
    return out


def arg_ordinals(cuda, kernel):
    L = cuda._L()
    ords = []
    for i in range(kernel.info.n_args):
        o = C.c_int32()
        cuda.check(L.cc_kernel_arg_param(kernel.handle, i, C.byref(o)))
        ords.append(o.value)
    return ords


_SIG = re.compile(r'extern "C" __global__ void(?: __launch_bounds__\([^)]*\))? (\w+)\(([^)]*)\)')


def _build(source: str, launches) -> C.CDLL:
    sigs = {m.group(1): [p.strip() for p in m.group(2).split(",") if p.strip()] for m in _SIG.finditer(source)}
    text = '#define CC_HOST_EMULATION 1\n#include "cuda_on_cpu.h"\n' + source + "\n"
    for i, li in enumerate(launches):
        entry = li.entry.decode()
        params = sigs[entry]
        assert len(params) == li.n_args, (entry, params, li.n_args)
        casts = []
        for j, prm in enumerate(params):
            ty = prm.replace("__restrict__", "").rsplit(None, 1)[0].strip()  # drop the parameter name
            casts.append(f"({ty})a[{j}]")
        g, b = list(li.grid), list(li.block)
        text += (f'extern "C" void emu_launch_{i}(void** a) {{ const unsigned g[3] = {{{g[0]}u, {g[1]}u, {g[2]}u}}, b[3] = {{{b[0]}u, {b[1]}u, {b[2]}u}};\n'
                 f"  emu::run_grid(g, b, [&] {{ {entry}({', '.join(casts)}); }}); }}\n")
    os.makedirs(_CACHE, exist_ok=True)
    key = hashlib.sha1((text + open(os.path.join(HERE, "cuda_on_cpu.h")).read()).encode()).hexdigest()[:20]
    so = os.path.join(_CACHE, key + ".so")
    if not os.path.exists(so):
        cpp = os.path.join(_CACHE, key + ".cpp")
        with open(cpp, "w") as f:
            f.write(text)
        cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-w", "-I", HERE, cpp, "-o", so + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host compile of the generated kernel failed:\n" + r.stderr[:4000])
        os.replace(so + ".tmp", so)
    return C.CDLL(so)


def _call(lib, i, li, args, out, scratch, partials, counter):
    ptrs = (C.c_void_p * li.n_args)()
    for j in range(li.n_args):
        a = li.args[j]
        if a >= 0:
            ptrs[j] = args[a].ctypes.data
        elif a == ARG_OUT:
            ptrs[j] = out.ctypes.data
        elif a == ARG_PARTIALS:
            ptrs[j] = partials.ctypes.data
        elif a == ARG_COUNTER:
            ptrs[j] = counter.ctypes.data
        elif a == ARG_COL_COUNTERS:
            ptrs[j] = COL_COUNTERS.ctypes.data
        elif a in (ARG_PEER_MB, ARG_PEER_EPOCH):
            ptrs[j] = 0  # no communicator on the host: the reduction stores its own values (mb_ == nullptr)
        else:
            ptrs[j] = scratch[ARG_SCRATCH0 - a].ctypes.data
    fn = getattr(lib, f"emu_launch_{i}")
    fn.argtypes = [C.POINTER(C.c_void_p)]
    fn.restype = None
    fn(ptrs)


def emulate(cuda, expr, leaf_arrays, max_threads=1 << 16):
    """Runs the plan `expr` compiles to on host threads. leaf_arrays[j] = data of the plan's j-th buffer argument.
    A contraction over gathered operand panels (kind 2 with generated panel kernels) runs its generated kernels here -- the panel
    gathers with their bounds tests / padding / TF32 split, and the in-place epilogue -- around a float64 product of the hi + lo panels
    standing in for the tcgen05 pipeline (which the GPU tier tests on its own); the precompiled plain contraction is not emulated."""
    k = expr.compile()
    info = k.info
    launches = launches_of(cuda, k)
    assert info.kind != 2 or launches, "the precompiled tensor-core contraction is not emulated"
    total = sum(int(np.prod(list(li.grid))) * int(np.prod(list(li.block))) for li in launches)
    assert total <= max_threads, f"{total} CUDA threads: too large for the host emulation"
    lib = _build(k.source, launches)
    args = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1).copy() for a in leaf_arrays]
    assert len(args) == info.n_args
    out = np.full(max(int(info.out_floats), 1) + 16, np.float32(-12345.0), np.float32)  # guard words past the end
    scratch = [np.zeros(int(n) + 16, np.float32) for n in list(launches[0].scratch_floats)[: launches[0].n_scratch]] if launches else []
    for sbuf in scratch:
        sbuf[-16:] = np.float32(-54321.0)
    partials = np.zeros(FOLD_PARTIALS + 16, np.float32)
    counter = np.zeros(4, np.uint32)
    n = int(info.out_floats)
    if info.kind == 2:
        m = re.search(r"general contraction (\d+)x(\d+)x(\d+) over gathered operand panels", k.source)
        assert m, "not a gathered-panel contraction"
        M, N, K = (int(g) for g in m.groups())
        mb = re.search(r"\(batch of (\d+)\)", k.source)
        batch = int(mb.group(1)) if mb else 1
        Kp = (K + 31) // 32 * 32
        assert [li.entry.decode() for li in launches[:2]] == ["panel_a", "panel_b"] and batch * M * N == n
        _call(lib, 0, launches[0], args, out, scratch, partials, counter)
        _call(lib, 1, launches[1], args, out, scratch, partials, counter)
        a_hi, a_lo, b_hi, b_lo = (sc[: batch * rows * Kp].reshape(batch * rows, Kp).astype(np.float64) for sc, rows in zip(scratch, (M, M, N, N)))
        for sc in scratch:
            assert (sc[-16:] == np.float32(-54321.0)).all(), "a panel kernel wrote past its panel"
        assert not a_hi[:, K:].any() and not a_lo[:, K:].any() and not b_hi[:, K:].any() and not b_lo[:, K:].any(), "K padding must be zero"
        for hi in (a_hi, b_hi):  # hi parts are exactly TF32 (10 explicit mantissa bits)
            assert not (hi.astype(np.float32).view(np.uint32) & 0x1FFF).any()
        a, bt = (a_hi + a_lo).reshape(batch, M, Kp), (b_hi + b_lo).reshape(batch, N, Kp)  # batch b: its own block of rows in each panel
        out[:n] = np.einsum("bmk,bnk->bmn", a, bt).astype(np.float32).reshape(-1)
        for i in range(2, len(launches)):
            _call(lib, i, launches[i], args, out, scratch, partials, counter)
    else:
        for i, li in enumerate(launches):
            _call(lib, i, li, args, out, scratch, partials, counter)
    assert (out[n:] == np.float32(-12345.0)).all(), "the kernel wrote past its output"
    assert not COL_COUNTERS.any(), "a fused second stage left its block counter set"
    return out[:n], k
