"""8192^3 contraction on the same kernel with different operand data (the tensor pipe is power-limited: throughput follows the bit activity of
the operands), with the SM clock and board power sampled while it runs"""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from compute.scala_b200 import cuda
cuda.init(0, streams=1)
T = cuda.Tensor
m = k = n = 8192
out = {}
for label, make in (("uniform [0,1) (Tensor.random)", lambda s: T.random([m, k], seed=s)), ("standard normal (Tensor.randomNormal)", lambda s: T.randomNormal([m, k], seed=s + (1 << 30))),
                    ("small integers {-4..4}", lambda s: ((T.random([m, k], seed=s) * T.fill(9.0, [m, k])) - (T.random([m, k], seed=s) * T.fill(9.0, [m, k])) % T.fill(1.0, [m, k]) - T.fill(4.0, [m, k]))),
                    ("zeros", lambda s: T.fill(0.0, [m, k]) + T.fill(0.0, [m, k]))):
    A, B = make(9).doCache(), make(10).doCache()
    a, b, c = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(m * n)
    cuda.set_operand_cache(False)
    for _ in range(3): cuda.matmul_3xtf32(a, b, c, m, n, k)
    cuda.synchronize()
    samples, stop = [], False
    def poll():
        while not stop:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
            try:
                f = r.stdout.strip().split(",")
                samples.append((float(f[0]), float(f[1])))
            except Exception:
                pass
            time.sleep(0.05)
    th = threading.Thread(target=poll); th.start()
    cuda.timer_start()
    steps = 60
    for _ in range(steps): cuda.matmul_3xtf32(a, b, c, m, n, k)
    ms = cuda.timer_stop() / steps
    stop = True; th.join()
    samples.sort()
    out[label] = {"ms": ms, "tflops": 2 * m * n * k / ms / 1e9, "sm_mhz_median": samples[len(samples) // 2][0] if samples else None,
                  "power_w_max": max(s[1] for s in samples) if samples else None}
    print(label, out[label], flush=True)
    for x in (a, b, c): x.release()
    del A, B
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gemm_data_power.json"), "w"), indent=1)
