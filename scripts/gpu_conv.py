"""The reference's Convolution benchmark (benchmarks.scala:412-622: 3x3 and 1x1 kernels, 32x32 images, batch 128 / 32, depth 8 / 3)
through the cuda backend: plan, kernel time, host time of building + compiling the graph. Run on the GPU box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0, streams=1)
T = cuda.Tensor


def convolute(inp, weight, bias):
    batch, height, width, depth = inp.shape
    kh, kw, _, filters = weight.shape
    input_seq = inp.split(3)
    outs = []
    bias_seq = bias.split(0)
    for f, khkwd in enumerate(weight.split(3)):
        summands = []
        for oy, kwd in zip(range(-(kh // 2), kh // 2 + 1), khkwd.split(0)):
            for ox, d in zip(range(-(kw // 2), kw // 2 + 1), kwd.split(0)):
                for in_c, w_c in zip(input_seq, d.split(0)):
                    summands.append(in_c.translate([0, oy, ox]) * w_c.broadcast([batch, height, width]))
        acc = summands[0]
        for s in summands[1:]:
            acc = acc + s
        outs.append(bias_seq[f].broadcast([batch, height, width]) + acc)
    return T.join(outs)


out = {}
cfgs = [(128, 32, 32, 8, 3), (32, 32, 32, 8, 3), (128, 32, 32, 3, 3), (128, 32, 32, 8, 1), (128, 32, 32, 32, 3)] + ([(64, 56, 56, 64, 3)] if "--big" in sys.argv else [])
for (b, h, w, d, ks) in cfgs:
    inp = T.randomNormal([b, h, w, d], seed=1).doCache()
    wt = T.randomNormal([ks, ks, d, d], seed=2).doCache()
    bias = T.randomNormal([d], seed=3).doCache()
    t0 = time.perf_counter()
    e = convolute(inp, wt, bias)
    t1 = time.perf_counter()
    k = e.compile()
    t2 = time.perf_counter()
    info = k.info
    for _ in range(3):
        e.doBuffer().release()
    cuda.synchronize()
    steps = 20
    cuda.timer_start()
    for _ in range(steps):
        e.doBuffer().release()
    ms = cuda.timer_stop() / steps
    flops = 2 * b * h * w * d * d * ks * ks
    name = f"batch{b} {h}x{w} depth{d} kernel{ks}x{ks}"
    out[name] = {"plan": info.kind, "ms": ms, "gflops": flops / ms / 1e6, "alg_GBs": info.algorithmic_bytes / ms / 1e6, "graph_build_s": t1 - t0,
                 "compile_s": t2 - t1, "note": k.source[:160].split("\n")[0]}
    print(name, out[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/conv.json", "w"), indent=1)
