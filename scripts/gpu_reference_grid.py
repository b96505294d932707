"""The reference's whole JMH parameter grid (benchmarks.scala:62-632: Issue137, MatrixMultiplication, Tanh, Sum, RandomNormal,
Convolution) through the cuda backend's public Tensor API, the way the JMH harness calls it: build the lazy graph, `flatArray`
(kernel + device->host read-back + wait), repeat.  Reports per-call wall time (JMH's throughput mode is 1 / this), the device
time of the same call, the plan the code generator chose, graph-build and JIT time of the first call, and a parity check of
every cell against the numpy oracle (tests-only code; this script is measurement tooling, not the product path).

    python scripts/gpu_reference_grid.py                 # on the GPU box -> gpurun_out/reference_grid.json
    python scripts/gpu_reference_grid.py --compile-only  # no GPU: graph build + plan + NVRTC time of every cell
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute.scala_b200 import cuda  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--compile-only", action="store_true")
ap.add_argument("--calls", type=int, default=30)
ap.add_argument("--no-parity", action="store_true")
args = ap.parse_args()

if not args.compile_only:
    cuda.init(0)
if not args.no_parity:
    from oracle import reference as ref  # noqa: E402

T = cuda.Tensor


# randomNormal's pair 61 ^ seed is (+inf, NaN) by construction (hash(61) == 0, Tensors.scala:106-117, 398-429); seeds with bit 30
# set put that pair beyond every size of the grid so the parity columns are about arithmetic, not about where the NaN lands
SEED = 1 << 30


def fold(n, x, f):
    for _ in range(n):
        x = f(x)
    return x


def matmul1(T_, m1, m2):  # benchmarks.scala:176-187 (unrolled when j*k is small)
    cols1 = m1.split(1)
    out = []
    for col2 in m2.split(1):
        acc = None
        for c1, s in zip(cols1, col2.split(0)):
            term = c1 * s.broadcast([m1.shape[0]])
            acc = term if acc is None else acc + term
        out.append(acc)
    return T_.join(out)


def matmul2(T_, m1, m2):  # benchmarks.scala:188-191
    i, j = m1.shape
    _, k = m2.shape
    product = m1.broadcast([i, j, k]) * m2.reshape([1, j, k]).broadcast([i, j, k])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def convolute(T_, inp, weight, bias):  # benchmarks.scala:463-556
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
    return T_.join(outs)


def cells():
    """(benchmark class, parameter string, builder(T) -> lazy result tensor, tolerance kind)"""
    for iters in (100, 10, 1):
        for nd in (3, 2):
            for size in (128, 32):
                def b(T_, iters=iters, nd=nd, size=size):
                    a, bb, c = (T_.randomNormal([size] * nd, seed=SEED + s) for s in (1, 2, 3))
                    return fold(iters, a, lambda x: x * bb + c)
                yield "Issue137", f"iterations={iters} dims={nd} size={size}", b, "chain"
    for ind in (8, 32):
        for outd in (8, 32):
            for batch in (65536, 4096, 32):
                def b(T_, ind=ind, outd=outd, batch=batch):
                    w = T_.randomNormal([ind, outd], seed=SEED + 1)
                    x = T_.randomNormal([batch, ind], seed=SEED + 2)
                    # benchmarks.scala:172-192: j and k unrolled when i >= maxComputeUnits * 128 (148 SMs), else only j
                    return matmul1(T_, x, w) if batch >= 148 * 128 else matmul2(T_, x, w)
                yield "MatrixMultiplication", f"inputDepth={ind} outputDepth={outd} batchSize={batch}", b, "sum"
    for iters in (100, 10, 1):
        for nd in (2, 3):
            for size in (128, 32):
                def b(T_, iters=iters, nd=nd, size=size):
                    return fold(iters, T_.randomNormal([size] * nd, seed=SEED + 1), T_.tanh)
                yield "Tanh", f"iterations={iters} dims={nd} size={size}", b, "ulp"
    for nd in (3, 2):
        for size in (512, 128, 32, 16):
            def b(T_, nd=nd, size=size):
                return T_.randomNormal([size] * nd, seed=SEED + 1).sum()
            yield "Sum", f"dims={nd} size={size}", b, "sum"
    for nd in (3, 2, 1):
        for size in (128, 32, 16):
            def b(T_, nd=nd, size=size):
                return T_.randomNormal([size] * nd, seed=SEED + 7)
            yield "RandomNormal", f"dims={nd} size={size}", b, "ulp"
    for ks in (3, 1):
        for depth in (8, 3):
            for batch in (128, 32):
                def b(T_, ks=ks, depth=depth, batch=batch):
                    inp = T_.randomNormal([batch, 32, 32, depth], seed=SEED + 1)
                    wt = T_.randomNormal([ks, ks, depth, depth], seed=SEED + 2)
                    bias = T_.randomNormal([depth], seed=SEED + 3)
                    return convolute(T_, inp, wt, bias)
                yield "Convolution", f"kernel={ks} depth={depth} batch={batch} image=32x32", b, "sum"


LEAVES = {}  # (shape, seed) -> the values the device generated, so that the oracle evaluates the SAME inputs


def cache_leaves(build):
    """JMH setup caches the inputs (`doCache`) and measures only the expression: make every randomNormal a cached tensor."""
    class Cached:
        def __getattr__(self, name):
            return getattr(T, name)

        @staticmethod
        def randomNormal(shape, seed):
            t = T.randomNormal(shape, seed=seed)
            if args.compile_only:
                return t
            t = t.doCache()
            if not args.no_parity:
                LEAVES[(tuple(shape), seed)] = t.flatArray()
            return t
    return build(Cached())


class OracleOnDeviceLeaves:
    """the numpy oracle's Tensor with randomNormal replaced by the device's values (the RNG's own parity is the RandomNormal rows:
    its log / cos / sin come from a different libm, <= 3 ulp apart, which would otherwise leak into every other row)"""

    def __getattr__(self, name):
        return getattr(ref.Tensor, name)

    @staticmethod
    def randomNormal(shape, seed):
        return ref.Tensor(LEAVES[(tuple(shape), seed)].reshape(shape))


out = {}
for klass, params, build, tol in cells():
    t0 = time.perf_counter()
    e = build(T) if klass == "RandomNormal" else cache_leaves(build)
    t1 = time.perf_counter()
    row = {"graph_build_ms": (t1 - t0) * 1e3}
    if klass != "RandomNormal":
        k = e.compile()
        t2 = time.perf_counter()
        info = k.info
        row.update(plan=int(info.kind), jit_ms=(t2 - t1) * 1e3, kernel_args=int(info.n_args))
    if not args.compile_only:
        got = e.flatArray()  # first call
        for _ in range(3):
            e.flatArray()
        t3 = time.perf_counter()
        for _ in range(args.calls):
            if klass == "RandomNormal":  # the JMH body builds a fresh tensor each call (benchmarks.scala:378-410)
                build(T).flatArray()
            else:
                e.flatArray()
        wall_us = (time.perf_counter() - t3) / args.calls * 1e6
        cuda.synchronize()
        cuda.timer_start()
        for _ in range(args.calls):
            e.doBuffer().release()
        dev_us = cuda.timer_stop() / args.calls * 1e3
        row.update(flatArray_wall_us=wall_us, device_us=dev_us, ops_per_s=1e6 / wall_us, elements=int(got.size))
        t4 = time.perf_counter()
        for _ in range(args.calls):
            (build(T) if klass == "RandomNormal" else e).flatBuffer().release()
        row["flatBuffer_wall_us"] = (time.perf_counter() - t4) / args.calls * 1e6
        if not args.no_parity:
            wants = []
            for contract in (False, True):  # the reference builds with -cl-unsafe-math-optimizations: a*b+c may or may not fuse
                ref.CONTRACT[0] = contract
                try:
                    wants.append(build(ref.Tensor if klass == "RandomNormal" else OracleOnDeviceLeaves()).flat_array())
                finally:
                    ref.CONTRACT[0] = False
            g64 = got.astype(np.float64)
            fin = np.isfinite(got)
            row["nonfinite_agree"] = bool(all(np.array_equal(np.isfinite(w), fin) for w in wants))
            if tol == "ulp":
                row["max_ulp_vs_oracle"] = int(min(ref.ulp_distance(got[fin], w[fin]).max() for w in wants))
            elif tol == "chain":
                # x -> x*b + c folded n times: errors of earlier steps are multiplied by later b's and cancellation in the last
                # add inflates them relative to the result; report the distribution of the distance to the nearer bracket end
                d = np.minimum(*[ref.ulp_distance(got[fin], w[fin]) for w in wants]).astype(np.float64)
                row["ulp_vs_oracle"] = {"median": float(np.median(d)), "p99": float(np.percentile(d, 99)), "max": float(d.max())}
            else:
                scale = float(np.abs(wants[0][fin]).max()) or 1.0
                row["max_abs_err_over_max"] = float(min(np.abs(g64[fin] - w.astype(np.float64)[fin]).max() for w in wants) / scale)
    LEAVES.clear()
    out.setdefault(klass, {})[params] = row
    print(klass, params, json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/reference_grid.json", "w"), indent=1)
