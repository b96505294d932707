"""A/B of the opt-in lowerings on a GPU (each arm in its own process: the switches are read at plan time / once):
  python scripts/gpu_knob_ab.py fuse_col     CC_FUSE_COL_STAGE      split axis reductions: second stage inside reduce_cols
  python scripts/gpu_knob_ab.py batched      CC_BATCHED_CONTRACTION batched matmul on the tcgen05 pipeline (one launch per batch)
  python scripts/gpu_knob_ab.py pdl          CC_PDL (default on)    programmatic dependent launch
  python scripts/gpu_knob_ab.py tile_owner   CC_REDUCE_TILE_OWNER   re-rolled reductions with a small trailing output dimension: a thread owns all of it
  python scripts/gpu_knob_ab.py small_n      CC_SMALL_N_MMA         small-N contractions (the reference's convolution benchmark sizes) on warp-level MMAs
Per workload: device time per step (CUDA events around the loop, best of 3), max |a - b| between the arms relative to max |a|.
Writes gpurun_out/knob_<name>.json.  (scripts/gpu_pdl.py is the earlier, PDL-only version with host submission times.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KNOBS = {"fuse_col": ("CC_FUSE_COL_STAGE", "0", "1"), "batched": ("CC_BATCHED_CONTRACTION", "0", "1"), "pdl": ("CC_PDL", "0", "1"),
         "tile_owner": ("CC_REDUCE_TILE_OWNER", "0", "1"),
         # off = the generic re-rolled reduction as it stood before both small-N lowerings
         "small_n": ("CC_SMALL_N_MMA", "0", "1", {"CC_REDUCE_TILE_OWNER": "0"})}


def chain(parts):
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def workloads(name, T):
    if name == "fuse_col":
        x1, x2, x3 = (T.random(s, seed=5).doCache() for s in ([16384, 16384], [16384, 4096], [4096, 4096]))
        x4 = T.random([64, 512, 256], seed=6).doCache()
        return [("C3 axis-0 sum 16384^2", lambda: chain(x1.split(0)), 100), ("axis-0 sum 16384x4096", lambda: chain(x2.split(0)), 200),
                ("axis-0 sum 4096^2", lambda: chain(x3.split(0)), 500), ("axis-0 sum 64x512x256", lambda: chain(x4.split(0)), 500),
                ("axis-0 sum 4096^2 + epilogue", lambda: T.tanh(chain(x3.split(0))) * T.fill(2.0, [4096]), 500)]
    if name == "batched":
        out = []
        for b, m, k, n in ((8, 512, 512, 512), (32, 256, 256, 256), (4, 2048, 1024, 2048), (64, 128, 64, 128)):
            A, B = T.randomNormal([b, m, k], seed=9).doCache(), T.randomNormal([b, k, n], seed=10).doCache()

            def build(A=A, B=B, b=b, m=m, k=k, n=n):
                return chain((A.broadcast([b, m, k, n]) * B.reshape([b, 1, k, n]).broadcast([b, m, k, n])).split(2))
            out.append((f"batched matmul {b}x{m}x{k}x{n}", build, 20))
        return out
    if name.startswith("red_p") or name in ("tile_owner", "small_n"):
        def conv(batch, size, depth, filters, ks=3):  # benchmarks.scala:463-556
            x, w, bias = T.randomNormal([batch, size, size, depth], seed=1).doCache(), T.randomNormal([ks, ks, depth, filters], seed=2).doCache(), T.randomNormal([filters], seed=3).doCache()
            xs, bs = x.split(3), bias.split(0)
            ws = [[[wc.split(0) for wc in wx.split(0)] for wx in wy.split(0)] for wy in w.split(0)]

            def build():
                outs = []
                for f in range(filters):
                    terms = [xs[c].translate([0, dy - ks // 2, dx - ks // 2]) * ws[dy][dx][c][f].broadcast([batch, size, size]) for dy in range(ks) for dx in range(ks) for c in range(depth)]
                    outs.append(bs[f].broadcast([batch, size, size]) + chain(terms))
                return T.join(outs)
            return build
    if name in ("tile_owner", "small_n"):
        # the reference's own benchmark grid (benchmarks.scala:612-630: kernel 3 / 1, batch 128 / 32, 32 x 32 images, depth 8 / 3) and its skinny products
        out = [(f"conv {ks}x{ks} batch {b} 32x32 depth {d}", conv(b, 32, d, d, ks), 500) for ks in (3, 1) for b in (128, 32) for d in (8, 3)]
        out.append(("conv 3x3 batch 64 64x64 depth 16", conv(64, 64, 16, 16), 100))
        for m, k, n in ((65536, 32, 32), (65536, 8, 8), (8192, 64, 16)):
            A, B = T.randomNormal([m, k], seed=4).doCache(), T.randomNormal([k, n], seed=5).doCache()
            out.append((f"matmul {m}x{k}x{n} (split / broadcast / sum)", lambda A=A, B=B, m=m, k=k, n=n: chain((A.broadcast([m, k, n]) * B.reshape([1, k, n]).broadcast([m, k, n])).split(1)), 500))
        # not a product of two loads: the tile owner's own territory (distances of many points to a few centres)
        for m, k, f in ((1 << 20, 16, 8), (1 << 18, 32, 16)):
            X, Cn = T.randomNormal([m, k], seed=6).doCache(), T.randomNormal([k, f], seed=7).doCache()
            out.append((f"L1 distances {m} points x {k} dims -> {f} centres", lambda X=X, Cn=Cn, m=m, k=k, f=f: chain(T.abs(X.broadcast([m, k, f]) - Cn.reshape([1, k, f]).broadcast([m, k, f])).split(1)), 200))
        return out
    if name.startswith("red_p"):
        A, B = T.random([512, 64], seed=4).doCache(), T.random([64, 512], seed=5).doCache()
        return [("conv 3x3 batch 128 32x32 depth 8", conv(128, 32, 8, 8), 500), ("conv 3x3 batch 32 32x32 depth 3", conv(32, 32, 3, 3), 1000),
                ("conv 3x3 batch 64 64x64 depth 16", conv(64, 64, 16, 16), 100),
                ("matmul 512x64x512 (generic reduction)", lambda: chain((A.broadcast([512, 64, 512]) * B.reshape([1, 64, 512]).broadcast([512, 64, 512])).split(1)), 500)]
    a, b, c = (T.random([1024, 1024], seed=s).doCache() for s in (1, 2, 3))
    x = T.random([4096, 4096], seed=5).doCache()
    return [("C1 tanh(a*b+c) 1024^2", lambda: T.tanh(a * b + c), 3000), ("axis-0 sum 4096^2", lambda: chain(x.split(0)), 500),
            ("axis-1 sum 4096^2", lambda: chain(x.split(1)), 500), ("fold of a*b 1024^2", lambda: (a * b).sum(), 2000)]


def arm(name):
    import numpy as np

    from compute.scala_b200 import cuda

    cuda.init(0)
    res = {}
    for label, build, steps in workloads(name, cuda.Tensor):
        e = build()
        k = e.compile()
        val = e.flatArray()
        np.save(os.path.join(ROOT, "gpurun_out", f"_knob_{os.environ.get('ARM', 'x')}_{len(res)}.npy"), val[: 1 << 20])
        best = None
        for _ in range(3):
            for _ in range(max(3, steps // 10)):
                e.doBuffer().release()
            cuda.synchronize()
            cuda.timer_start()
            for _ in range(steps):
                e.doBuffer().release()
            ms = cuda.timer_stop()
            best = ms if best is None else min(best, ms)
        res[label] = {"us_per_step": best / steps * 1e3, "plan": k.info.kind, "launches": k.info.n_launches, "note": k.source.split("\n", 1)[0][:200]}
    print(json.dumps(res))


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[2] == "arm":
        arm(sys.argv[1])
        sys.exit(0)
    import numpy as np

    name = sys.argv[1]
    env_name, off, on = KNOBS[name][:3]
    off_extra = KNOBS[name][3] if len(KNOBS[name]) > 3 else {}
    res = {}
    for tag, v in (("off", off), ("on", on)):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), name, "arm"], env=dict(os.environ, ARM=tag, **{env_name: v}, **(off_extra if tag == "off" else {})),
                           capture_output=True, text=True, timeout=900)
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-1500:]}
    if all("error" not in v for v in res.values()):
        res["speedup"], res["max_rel_diff"] = {}, {}
        for i, label in enumerate(res["off"]):
            res["speedup"][label] = res["off"][label]["us_per_step"] / res["on"][label]["us_per_step"]
            a, b = (np.load(os.path.join(ROOT, "gpurun_out", f"_knob_{t}_{i}.npy")) for t in ("off", "on"))
            res["max_rel_diff"][label] = float(np.abs(a - b).max() / max(1e-30, float(np.abs(a).max())))
            res.setdefault("bit_identical", {})[label] = bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    for f in os.listdir(os.path.join(ROOT, "gpurun_out")):
        if f.startswith("_knob_"):
            os.remove(os.path.join(ROOT, "gpurun_out", f))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"knob_{name}.json"), "w"), indent=1)
    print(json.dumps(res))
