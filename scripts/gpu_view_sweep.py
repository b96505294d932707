"""HBM throughput of the view / reduction templates away from the BASELINE shapes: misaligned translations, odd extents,
middle-axis permutes and reductions, fused operands.  One line per pattern: plan, ms, algorithmic GB/s, fraction of the
measured HBM peak.  Run on the GPU box; results land in gpurun_out/view_sweep.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0, streams=1)
T = cuda.Tensor
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6549.8


def chain_sum(parts):
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def n_of(shape):
    n = 1
    for s in shape:
        n *= s
    return n


cases = []


def case(name, build, bytes_moved):
    cases.append((name, build, bytes_moved))


N3 = [512, 512, 512]
B3 = 4 * n_of(N3)
case("identity copy 512^3 (x + 0 literal)", lambda: X3 + T.fill(0.0, N3), 2 * B3)
case("translate [0,0,1] 512^3 (misaligned by one float)", lambda: X3.translate([0, 0, 1]), 2 * B3)
case("translate [0,0,4] 512^3 (aligned)", lambda: X3.translate([0, 0, 4]), 2 * B3)
case("translate [1,-2,3] 512^3", lambda: X3.translate([1, -2, 3]), 2 * B3)
case("x + x.translate([0,0,1]) 512^3 (stencil, second read hits L2/L1)", lambda: X3 + X3.translate([0, 0, 1]), 2 * B3)
case("5-point stencil along the last two dims 512^3", lambda: X3 + X3.translate([0, 0, 1]) + X3.translate([0, 0, -1]) + X3.translate([0, 1, 0]) + X3.translate([0, -1, 0]), 2 * B3)
case("permute [1,0,2] 512^3 (rows moved whole)", lambda: X3.permute([1, 0, 2]), 2 * B3)
case("permute [0,2,1] 512^3 (batched transpose)", lambda: X3.permute([0, 2, 1]), 2 * B3)
case("permute [2,1,0] 512^3", lambda: X3.permute([2, 1, 0]), 2 * B3)
case("permute [2,0,1] + translate [3,-5,7] 512^3 (C4)", lambda: X3.permute([2, 0, 1]).translate([3, -5, 7]), 2 * B3)
case("tanh(x.permute([2,1,0]) * y) 512^3", lambda: T.tanh(X3.permute([2, 1, 0]) * Y3), 3 * B3)
OD = [1001, 1003, 127]
case("odd extents 1001x1003x127: x * y (flat)", lambda: XO * YO, 3 * 4 * n_of(OD))
case("odd extents 1001x1003x127: translate [0,0,1]", lambda: XO.translate([0, 0, 1]), 2 * 4 * n_of(OD))
case("odd extents 1001x1003x127: permute [2,0,1]", lambda: XO.permute([2, 0, 1]), 2 * 4 * n_of(OD))
case("odd extents 1001x1003x127: permute [1,0,2]", lambda: XO.permute([1, 0, 2]), 2 * 4 * n_of(OD))
case("transpose 8192x8200", lambda: XT.transpose(), 2 * 4 * 8192 * 8200)
case("transpose 8192x8191 (odd)", lambda: XT2.transpose(), 2 * 4 * 8192 * 8191)
case("broadcast row [512] -> 512^3 times x", lambda: X3 * R3, 2 * B3)
case("join(split(0)) along dim 0 512^3", lambda: T.join(X3.split(0), 0), 2 * B3)
case("join(split(2)) along dim 2 512^3 (last dim)", lambda: T.join(X3.split(2)), 2 * B3)
S2 = [4096, 4096]


def window(x, f):
    terms = [x.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    return chain_with(terms, f)


def chain_with(parts, f):
    acc = parts[0]
    for q in parts[1:]:
        acc = f(acc, q)
    return acc


case("3x3 box sum (9 translated views) 4096^2", lambda: window(XS, lambda a, b: a + b), 2 * 4 * 4096 * 4096)
case("3x3 max pool, stride 1 (9 translated views) 4096^2", lambda: window(XS, T.max), 2 * 4 * 4096 * 4096)
case("5x5 box sum (25 translated views) 4096^2", lambda: chain_with([XS.translate([dy, dx]) for dy in range(-2, 3) for dx in range(-2, 3)], lambda a, b: a + b), 2 * 4 * 4096 * 4096)
# the same windows where fixed costs (launch, the first tile's latency, the tail) weigh less, and the plain copy at both sizes for scale
case("identity copy 4096^2 (x + 0 literal)", lambda: XS + T.fill(0.0, S2), 2 * 4 * 4096 * 4096)
case("3x3 box sum (9 translated views) 16384x8192", lambda: chain_with([XL.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], lambda a, b: a + b), 2 * 4 * 16384 * 8192)
case("5x5 box sum (25 translated views) 16384x8192", lambda: chain_with([XL.translate([dy, dx]) for dy in range(-2, 3) for dx in range(-2, 3)], lambda a, b: a + b), 2 * 4 * 16384 * 8192)
case("identity copy 16384x8192 (x + 0 literal)", lambda: XL + T.fill(0.0, [16384, 8192]), 2 * 4 * 16384 * 8192)
M3 = [256, 512, 1024]
case("sum over middle axis 256x512x1024", lambda: chain_sum(XM.split(1)), 4 * n_of(M3) + 4 * 256 * 1024)
case("sum over last axis 256x512x1024", lambda: chain_sum(XM.split(2)), 4 * n_of(M3) + 4 * 256 * 512)
case("sum over first axis 256x512x1024", lambda: chain_sum(XM.split(0)), 4 * n_of(M3) + 4 * 512 * 1024)
case("sum over axis 0 of inline a*b 16384x4096 (operand fused)", lambda: chain_sum((XA * XB).split(0)), 2 * 4 * 16384 * 4096 + 4 * 4096)
case("sum over axis 1 of inline tanh(a)*b 16384x4096", lambda: chain_sum((T.tanh(XA) * XB).split(1)), 2 * 4 * 16384 * 4096 + 4 * 16384)
case("max over everything of a*b 16384x4096 (fold, fused)", lambda: (XA * XB).reduce("max"), 2 * 4 * 16384 * 4096)
case("depth-3 channels last 1024x1024x3: x * w.broadcast (odd fastest dim)", lambda: XC * WC3, 2 * 4 * 3 * 1024 * 1024)

X3 = T.random(N3, seed=1).doCache()
Y3 = T.random(N3, seed=2).doCache()
XO = T.random(OD, seed=3).doCache()
YO = T.random(OD, seed=4).doCache()
XT = T.random([8192, 8200], seed=5).doCache()
XT2 = T.random([8192, 8191], seed=6).doCache()
ROW = T.random([512], seed=7).doCache()
XM = T.random(M3, seed=8).doCache()
XS = T.random(S2, seed=13).doCache()
XL = T.random([16384, 8192], seed=14).doCache()
XA = T.random([16384, 4096], seed=9).doCache()
XB = T.random([16384, 4096], seed=10).doCache()
XC = T.random([1024, 1024, 3], seed=11).doCache()
WC = T.random([3], seed=12).doCache()
# leading-aligned broadcast (T:833-855): [512,512] -> 512^3 repeats along the LAST dim; a per-last-dim row needs a permute
R3 = T.random([512], seed=7).doCache().broadcast([512, 512, 512]).permute([2, 1, 0])  # value depends on the last index only
WC3 = WC.broadcast([3, 1024, 1024]).permute([1, 2, 0])

out = {}
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, build, nbytes in cases:
    if only and only not in name:
        continue
    try:
        e = build()
        k = e.compile()
        kind = int(k.info.kind)
        for _ in range(3):
            e.doBuffer().release()
        cuda.synchronize()
        steps = 10
        cuda.stats_reset()
        cuda.timer_start()
        for _ in range(steps):
            e.doBuffer().release()
        ms = cuda.timer_stop() / steps
        launches = cuda.stats()["device_kernels"] / steps
        row = {"plan": kind, "ms": ms, "kernels": launches, "GBs": nbytes / ms / 1e6, "frac_of_hbm": nbytes / ms / 1e6 / PEAK,
               "template": k.source[k.source.rfind("\n// ", 0, k.source.rfind("extern \"C\"")) + 1:].split("\n")[0][:140]}
        del e, k
    except Exception as ex:  # noqa: BLE001
        row = {"error": str(ex)[:300]}
    out[name] = row
    print(f"{name:75s} {json.dumps(row)}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/view_sweep.json", "w"), indent=1)
