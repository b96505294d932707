"""A/B of the contraction's tile configurations at large sizes: device time per call of cc_matmul_3xtf32 (B panels cached / not cached)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
out = {}
for (m, k, n) in [(8192, 8192, 8192), (4096, 4096, 4096), (2048, 2048, 2048), (1024, 1024, 1024), (1024, 8192, 8192), (65536, 32, 32)]:
    A, B = T.random([m, k], seed=9).doCache(), T.random([k, n], seed=10).doCache()
    a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
    rec = {}
    for cache in (False, True):
        cuda.set_operand_cache(cache)
        for _ in range(3): cuda.matmul_3xtf32(a, b, c, m, n, k)
        cuda.synchronize()
        steps = 10 if m * n * k >= 2**36 else 50
        cuda.timer_start()
        for _ in range(steps): cuda.matmul_3xtf32(a, b, c, m, n, k)
        ms = cuda.timer_stop() / steps
        rec["b_cached" if cache else "fresh"] = {"ms": ms, "tflops": 2 * m * n * k / ms / 1e9}
    out["%%dx%%dx%%d" %% (m, k, n)] = rec
    for x in (a, b, c): x.release()
print(json.dumps(out))
''' % ROOT
res = {}
for cfg in ("1024", "1025", "512"):
    env = dict(os.environ)
    if cfg: env["CC_GEMM_FORCE_CONFIG"] = cfg
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    try:
        res[cfg or "picker"] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[cfg or "picker"] = {"error": (r.stdout + r.stderr)[-600:]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_v2.json"), "w"), indent=1)
for cfg, v in res.items():
    print(cfg, {k: {a: round(b["tflops"], 1) for a, b in x.items()} for k, x in v.items()} if "error" not in v else v)
