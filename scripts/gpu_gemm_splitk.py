"""small products on the tensor-memory-A kernel: K split over the CTA pairs (CC_GEMM_K_SPLITS) against the launcher's own choice; per call
of cc_matmul_3xtf32: device time eager, device time replayed from a captured graph (no host cost between kernels), host wall time"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, time
sys.path.insert(0, %r)
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
out = {}
for (m, k, n) in [(1024, 1024, 1024), (512, 4096, 512), (1536, 1536, 1536), (2048, 2048, 2048), (1024, 8192, 1024), (768, 768, 768)]:
    A, B = T.random([m, k], seed=9).doCache(), T.random([k, n], seed=10).doCache()
    a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
    rec = {}
    for cache in (False, True):
        cuda.set_operand_cache(cache)
        for _ in range(3): cuda.matmul_3xtf32(a, b, c, m, n, k)
        cuda.synchronize()
        steps = 50
        t0 = time.perf_counter()
        cuda.timer_start()
        for _ in range(steps): cuda.matmul_3xtf32(a, b, c, m, n, k)
        host = (time.perf_counter() - t0) / steps
        ms = cuda.timer_stop() / steps
        with cuda.Graph() as g:
            for _ in range(20): cuda.matmul_3xtf32(a, b, c, m, n, k)
        g.launch(); cuda.synchronize()
        cuda.timer_start()
        for _ in range(5): g.launch()
        gms = cuda.timer_stop() / 100
        g.release()
        rec["b_cached" if cache else "fresh"] = {"eager_us": ms * 1e3, "graph_us": gms * 1e3, "host_us": host * 1e6, "graph_tflops": 2 * m * n * k / gms / 1e9}
    out["%%dx%%dx%%d" %% (m, k, n)] = rec
    for x in (a, b, c): x.release()
print(json.dumps(out))
''' % ROOT
res = {}
for label, env_add in (("auto", {}), ("splits=1", {"CC_GEMM_K_SPLITS": "1"}), ("splits=2", {"CC_GEMM_K_SPLITS": "2"}), ("splits=4", {"CC_GEMM_K_SPLITS": "4"}),
                       ("splits=8", {"CC_GEMM_K_SPLITS": "8"}), ("tmem_a_off", {"CC_GEMM_TMEM_A": "0"})):
    env = dict(os.environ, **env_add)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    try:
        res[label] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[label] = {"error": (r.stdout + r.stderr)[-600:]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_splitk.json"), "w"), indent=1)
for label, v in res.items():
    if "error" in v:
        print(label, v); continue
    for shape, x in v.items():
        print(label, shape, {kk: {a: round(bv, 1) for a, bv in vv.items()} for kk, vv in x.items()})
