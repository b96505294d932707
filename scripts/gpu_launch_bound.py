"""Is a launch-bound loop bound by the host or by the device?  Short bursts (fewer launches than the driver's queue holds) time the
host alone: submission never blocks, so (burst wall time) / n is the cost of one step on the host with the real driver; the same
burst plus the final synchronize, and a long loop, give the device's rate.  Writes gpurun_out/launch_bound.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0)
T = cuda.Tensor
out = {}
for name, n in (("C1 tanh(a*b+c) 1024^2", 1024), ("tanh(a*b+c) 32^2", 32)):
    a, b, c = (T.random([n, n], seed=s).doCache() for s in (1, 2, 3))
    e = T.tanh(a * b + c)
    e.flatArray()
    step = lambda: e.doBuffer().release()  # noqa: E731
    for _ in range(2000):
        step()
    cuda.synchronize()
    rec = {}
    for burst in (50, 200):
        host, total = [], []
        for _ in range(30):
            cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(burst):
                step()
            t1 = time.perf_counter()
            cuda.synchronize()
            t2 = time.perf_counter()
            host.append((t1 - t0) / burst * 1e6)
            total.append((t2 - t0) / burst * 1e6)
        host.sort(), total.sort()
        rec[f"burst_{burst}"] = {"host_us_per_step_median": host[len(host) // 2], "host_us_per_step_min": host[0],
                                 "burst_plus_sync_us_per_step_median": total[len(total) // 2]}
    cuda.timer_start()
    t0 = time.perf_counter()
    for _ in range(20000):
        step()
    t1 = time.perf_counter()
    ms = cuda.timer_stop()
    rec["long_loop"] = {"device_us_per_step": ms / 20000 * 1e3, "host_us_per_step": (t1 - t0) / 20000 * 1e6}
    out[name] = rec
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "launch_bound.json"), "w"), indent=1)
print(json.dumps(out))
