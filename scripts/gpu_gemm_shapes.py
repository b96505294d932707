"""Contraction (tcgen05 3xTF32) vs the generic JIT reduction for the matmul pattern over a range of shapes, and the fused
whole-tensor fold vs materialise-then-sum. Run on the GPU box: python scripts/gpu_gemm_shapes.py [generic]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
generic = len(sys.argv) > 1 and sys.argv[1] == "generic"
if generic:
    os.environ["CC_DISABLE_CONTRACTION"] = "1"
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0, streams=1)
T = cuda.Tensor


def matmul2(a, b):
    i, j = a.shape
    _, k = b.shape
    p = a.broadcast([i, j, k]) * b.reshape([1, j, k]).broadcast([i, j, k])
    parts = p.split(1)
    acc = parts[0]
    for q in parts[1:]:
        acc = acc + q
    return acc


def timeit(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    cuda.synchronize()
    cuda.timer_start()
    for _ in range(steps):
        fn()
    return cuda.timer_stop() / steps


shapes = [(128, 128, 128), (256, 256, 256), (512, 512, 512), (1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (1000, 1000, 1000),
          (65536, 32, 32), (65536, 8, 8), (4096, 32, 32), (4096, 256, 256), (8192, 64, 64), (8192, 128, 128), (300, 300, 2000), (8192, 8192, 8192)]
out = {}
for (m, k, n) in shapes:
    if generic and m * k * n > 2**32:
        continue
    A, B = T.randomNormal([m, k], seed=9).doCache(), T.randomNormal([k, n], seed=10).doCache()
    if generic:
        e = matmul2(A, B)
        kind = e.compile().info.kind
        ms = timeit(lambda: e.doBuffer().release(), steps=5 if m * k * n > 2**30 else 20)
    else:
        ab, bb, cb = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
        kind = 2
        ms = timeit(lambda: cuda.matmul_3xtf32(ab, bb, cb, m, n, k), steps=5 if m * k * n > 2**34 else 20)
        for x in (ab, bb, cb):
            x.release()
    out[f"{m}x{k}x{n}"] = {"ms": round(ms, 5), "tflops": round(2 * m * n * k / ms / 1e9, 2), "plan": kind}
    print(f"{m}x{k}x{n}", out[f"{m}x{k}x{n}"], flush=True)

if not generic:
    n = 16384
    a, b, c = (T.random([n, n], seed=s).doCache() for s in (1, 2, 3))
    t = a * b + c
    chain = T.tanh(T.log(T.exp(t) + a) * b) + c
    fused = chain.sum()
    ms_f = timeit(lambda: fused.doBuffer().release(), steps=20)
    ms_2 = timeit(lambda: chain.nonInline().sum().doBuffer().release(), steps=20)
    x = T.random([n, n], seed=5).doCache()
    ms_s = timeit(lambda: x.sum().doBuffer().release(), steps=20)
    ms_m = timeit(lambda: x.reduce("max").doBuffer().release(), steps=20)
    out["fold"] = {"fused chain.sum ms": ms_f, "GB/s (12 B/elt)": 12 * n * n / ms_f / 1e6, "materialise then sum ms": ms_2, "plain sum ms": ms_s,
                   "plain sum GB/s": 4 * n * n / ms_s / 1e6, "jit max ms": ms_m, "jit max GB/s": 4 * n * n / ms_m / 1e6}
    print(out["fold"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/gemm_shapes_{'generic' if generic else 'tcgen05'}.json", "w"), indent=1)
