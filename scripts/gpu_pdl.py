"""A/B of programmatic dependent launch (CC_PDL=1, runtime.cpp `pdl_enabled`): per-step device time of launch-bound configurations
with and without it (separate processes: the switch is read once), results compared bit for bit.
  python scripts/gpu_pdl.py            -> runs both arms, prints one JSON line, writes gpurun_out/pdl.json"""
import json
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def arm():
    import numpy as np

    from compute.scala_b200 import cuda

    cuda.init(0)
    T = cuda.Tensor
    out = {}

    def chain(parts):
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    def measure(name, build, steps, warmup):
        e = build()
        crc = zlib.crc32(np.ascontiguousarray(e.flatArray()).tobytes())

        def step():
            e.doBuffer().release()

        import time

        best = host = None
        for _ in range(3):
            for _ in range(warmup):
                step()
            cuda.synchronize()
            cuda.timer_start()
            t0 = time.perf_counter()
            for _ in range(steps):
                step()
            t1 = time.perf_counter()  # the host is done submitting; the device may still be running
            ms = cuda.timer_stop()
            best = ms if best is None else min(best, ms)
            host = (t1 - t0) if host is None else min(host, t1 - t0)
        out[name] = {"us_per_step": best / steps * 1e3, "host_submit_us_per_step": host / steps * 1e6, "crc32": crc}

    a, b, c = (T.random([1024, 1024], seed=s).doCache() for s in (1, 2, 3))
    measure("C1 tanh(a*b+c) 1024^2", lambda: T.tanh(a * b + c), 3000, 1000)
    s1, s2, s3 = (T.random([32, 32], seed=s).doCache() for s in (1, 2, 3))
    measure("tanh(a*b+c) 32^2", lambda: T.tanh(s1 * s2 + s3), 3000, 1000)
    x = T.random([4096, 4096], seed=5).doCache()
    measure("axis-0 sum 4096^2 (two launches)", lambda: chain(x.split(0)), 500, 100)
    measure("axis-1 sum 4096^2", lambda: chain(x.split(1)), 500, 100)
    measure("full sum of a*b 1024^2 (fold kernel)", lambda: (a * b).sum(), 2000, 500)
    measure("transpose 2048^2", lambda: T.random([2048, 2048], seed=9).doCache().transpose() + T.fill(1.0, [2048, 2048]), 1000, 200)
    big = T.random([8192, 8192], seed=4).doCache()
    measure("abs 8192^2 (many waves)", lambda: T.abs(big) * big, 200, 20)
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "arm":
        arm()
        sys.exit(0)
    res = {}
    for pdl in ("0", "1"):
        env = dict(os.environ, CC_PDL=pdl)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "arm"], env=env, capture_output=True, text=True, timeout=120)
        if r.returncode != 0:
            res["pdl" + pdl] = {"error": (r.stderr or r.stdout)[-1500:]}
        else:
            res["pdl" + pdl] = json.loads(r.stdout.strip().splitlines()[-1])
    if all("error" not in v for v in res.values()):
        res["same_results"] = all(res["pdl0"][k]["crc32"] == res["pdl1"][k]["crc32"] for k in res["pdl0"])
        res["speedup"] = {k: res["pdl0"][k]["us_per_step"] / res["pdl1"][k]["us_per_step"] for k in res["pdl0"]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "pdl.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
