#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configuration.

Headline line (one JSON object on stdout, rank 0): config C2 = the long fused elementwise chain
`out = tanh(log(exp(a*b+c)+a)*b)+c` on 2^28 float32 elements per GPU, inputs `random([16384,16384], seed 1,2,3)`
(Wang-hash uniform, Tensors.scala:432-443), evaluated through the lazy Tensor API -> C ABI -> one JIT-compiled sm_100a
kernel per step.  `value` = algorithmic bytes (16 B/element) / device time with inputs resident in HBM;
`e2e` = the same metric through the same API with HOST buffers (pinned), H2D of the three inputs and D2H of the result
inside the timed region.  Other configs (C1, C3, C4, C5) are measured briefly and reported under "configs".

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, weak scaling)
    python bench.py --impl reference        (CPU port of the reference's generated kernel on the host cores)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS = 16384, 16384
BYTES_PER_ELEMENT = 16  # 3 reads + 1 write of fp32
METRIC = "fused elementwise HBM GB/s (C2 chain, 2^28 fp32 elements per GPU)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0: float, t1: float) -> dict:
        good = [(t, r) for t, r in self.rows if len(r) >= 7]
        rows = [r for t, r in good if t0 <= t <= t1 + 0.03]
        if len(rows) < 3 and good:  # timed region shorter than a few sampling periods: take the samples nearest to it
            mid = 0.5 * (t0 + t1)
            rows = [r for _, r in sorted(good, key=lambda tr: abs(tr[0] - mid))[:5]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


def c2_chain(T, a, b, c):
    t = a * b + c
    u = T.exp(t)
    v = T.log(u + a)
    w = T.tanh(v * b)
    return w + c


def convolute(T, inp, weight, bias):
    """benchmarks.scala:463-556, written the way the reference writes it: split / translate / broadcast / join"""
    batch, height, width, depth = inp.shape
    kh, kw, _, filters = weight.shape
    input_seq = inp.split(3)
    bias_seq = bias.split(0)
    outs = []
    for f, khkwd in enumerate(weight.split(3)):
        summands = []
        for oy, kwd in zip(range(-(kh // 2), kh // 2 + 1), khkwd.split(0)):
            for ox, d in zip(range(-(kw // 2), kw // 2 + 1), kwd.split(0)):
                for in_c, w_c in zip(input_seq, d.split(0)):
                    summands.append(in_c.translate([0, oy, ox]) * w_c.broadcast([batch, height, width]))
        acc = summands[0]
        for x in summands[1:]:
            acc = acc + x
        outs.append(bias_seq[f].broadcast([batch, height, width]) + acc)
    return T.join(outs)


# ---- reference arm: the CPU port of the generated kernel on the host cores -----------------------------------------------


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_lib():
    """the timed CPU baseline: oracle/oracle_cpu.c built ON THIS MACHINE with -O3 -march=native -fopenmp -ffast-math (BASELINE.md section 4;
    -ffast-math is the host analogue of the reference's -cl-unsafe-math-optimizations and lets gcc call glibc's vector expf / logf / tanhf the
    way POCL's work-group vectoriser would). The thread count is set explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank."""
    from oracle import build as ob

    L = ob.load("native")
    L.oracle_set_num_threads(host_threads())
    return L


def cpu_c2(n: int, warmup: int, reps: int, budget_s: float | None = None):
    """reps timed passes of the generated C2 kernel over n elements on all host threads (after `warmup` untimed ones);
    with a budget, reps shrinks so that the timed passes take about that long (never below 3)"""
    L = cpu_lib()
    a, b, c, out = (np.empty(n, np.float32) for _ in range(4))
    for arr, seed in ((a, 1), (b, 2), (c, 3)):
        L.oracle_random(arr.ctypes.data, n, seed)
    t0 = time.perf_counter()
    L.oracle_c2(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n)  # first pass: page faults of `out`, libm set-up
    first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        L.oracle_c2(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n)
    if budget_s is not None:
        reps = int(max(3, min(reps, budget_s / max(first, 1e-3))))
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        L.oracle_c2(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n)
        times.append(time.perf_counter() - t0)
    # the fast-math build is the same kernel: its output stays within a few ulp-sized steps of the strict port's on a sample
    from oracle import build as ob

    m = min(n, 1 << 16)
    strict = np.empty(m, np.float32)
    ob.load("strict").oracle_c2(a.ctypes.data, b.ctypes.data, c.ctypes.data, strict.ctypes.data, m)
    drift = float(np.abs(out[:m] - strict).max())
    assert drift <= 5e-6, f"the -ffast-math CPU baseline drifted {drift} from the strict port"
    return times, int(L.oracle_num_threads()), out


C2_WORKLOAD = "C2 long fused elementwise chain tanh(log(exp(a*b+c)+a)*b)+c"
CPU_NOTE = ("the reference (Scala + LWJGL OpenCL on POCL) cannot run in this image (no JVM, no OpenCL ICD); this is the C/OpenMP port of the kernel it "
            "generates (oracle/oracle_cpu.c), gcc -O3 -march=native -fopenmp -ffast-math built on this host, all host threads")


def run_reference(args, rank: int):
    if rank != 0:
        return
    # each step = one pass of the generated kernel over the FULL per-GPU workload (2^28 elements, the cuda arm's config);
    # BENCH_REFERENCE_SAMPLE_LOG2 shrinks it for the CPU-only contract test
    log2n = int(os.environ.get("BENCH_REFERENCE_SAMPLE_LOG2", "28"))
    n = 1 << log2n
    times, cores, _ = cpu_c2(n, args.warmup, args.steps)
    sec = sum(times) / len(times)
    gbs = BYTES_PER_ELEMENT * n / sec / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C2_WORKLOAD, "shape_per_gpu": [n // COLS, COLS] if n >= COLS else [1, n], "elements_per_gpu": n,
                   "inputs": "Tensor.random seeds 1,2,3 (Wang hash), resident in host memory", "note": CPU_NOTE},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"all 2^{log2n} elements per step, {len(times)} timed steps after {args.warmup} warm-up passes"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_side_baselines() -> dict:
    """C3 and C5 on the host, the way the reference's `cpu` backend runs them (bounded samples; reported beside the GPU figures)"""
    L = cpu_lib()
    cores = int(L.oracle_num_threads())
    out = {}
    n = 1 << 28
    x = np.empty(n, np.float32)
    L.oracle_random(x.ctypes.data, n, 5)
    L.oracle_sum_cpu_order(x.ctypes.data, 1 << 20)
    t0 = time.perf_counter()
    L.oracle_sum_cpu_order(x.ctypes.data, n)
    sec = time.perf_counter() - t0
    out["C3 full sum 16384^2"] = {"value": 4 * n / sec / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
                                 "sample": "one pass over all 2^28 elements; ONE thread, as the reference launches its CPU reduction "
                                           "(global = local = 1, Tensors.scala:690-695)"}
    for axis in (0, 1):
        res = np.empty(ROWS, np.float32)
        t0 = time.perf_counter()
        L.oracle_axis_sum_2d(x.ctypes.data, ROWS, COLS, axis, res.ctypes.data)
        sec = time.perf_counter() - t0
        out[f"C3 axis-{axis} sum 16384^2"] = {"value": 4 * n / sec / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                              "sample": "one pass; one work-item per output element, fp32 left fold over the split index"}
    del x
    m = 1024
    a, b, c = (np.empty(m * m, np.float32) for _ in range(3))
    L.oracle_random(a.ctypes.data, m * m, 9)
    L.oracle_random(b.ctypes.data, m * m, 10)
    L.oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, m, m)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        L.oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, m, m)
    sec = (time.perf_counter() - t0) / reps
    out["C5 matmul as split/broadcast/sum"] = {"value": 2 * m**3 / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                                               "sample": f"1024^3 (the full 8192^3 cannot run on the reference at all: SURVEY finding 2), mean of {reps} passes, "
                                                         "fp32 left fold per output row"}
    return out


# ---- cuda arm ---------------------------------------------------------------------------------------------------------------------


def time_steps(cuda, fn, steps: int, warmup: int):
    for _ in range(warmup):
        fn()
    cuda.synchronize()
    s0 = cuda.stats()
    cuda.timer_start()
    w0 = time.time()
    for _ in range(steps):
        fn()
    ms = cuda.timer_stop()
    w1 = time.time()
    s1 = cuda.stats()
    return ms, s1["device_kernels"] - s0["device_kernels"], w0, w1


SHORT_SIDE = False


def side_configs(cuda, hbm_peak: float, tf_peak: float) -> dict:
    """C1, C3, C4 (and C5 when the contraction is built) — short device-resident measurements for the same JSON line"""
    T = cuda.Tensor
    out = {}

    def measure(name, build, alg_bytes, steps=10, flops=None, warmup=3, best_of=1):
        try:
            expr = build()
            k = expr.compile()
            kind = k.info.kind
            k.release()

            def step():
                expr.doBuffer().release()

            ms, launches, _, _ = time_steps(cuda, step, steps, warmup)
            windows = [ms / steps]
            for _ in range(best_of - 1):  # (a host-bound loop: other processes on the box show up as slow windows)
                ms, launches, _, _ = time_steps(cuda, step, steps, 0)
                windows.append(ms / steps)
            per = min(windows)
            rec = {"ms": per, "kernels_per_step": launches / steps, "plan": kind}
            if best_of > 1:
                rec["windows_ms"] = windows
                rec["note"] = f"best of {best_of} windows of {steps} steps (bound by the host's submission rate, not by the kernel)"
            if flops:
                rec["tflops"] = flops / per / 1e9
                rec["frac_of_3xtf32_peak"] = rec["tflops"] / tf_peak
            else:
                rec["gbs"] = alg_bytes / per / 1e6
                rec["frac_of_hbm"] = rec["gbs"] / hbm_peak
            out[name] = rec
        except Exception as e:  # a side measurement must not take the headline down
            out[name] = {"error": str(e)[:200]}

    n1 = 1024
    a1, b1, c1 = (T.random([n1, n1], seed=s).doCache() for s in (1, 2, 3))
    # a 5 us step: enough warm-up and steps that the clock ramp after the idle CPU-baseline phase is not what gets timed
    # (--short-side: a handful of steps only, so that an ncu launch list of the whole run stays short)
    n_c1 = 5 if SHORT_SIDE else 2000
    measure("C1 tanh(a*b+c) 1024^2", lambda: T.tanh(a1 * b1 + c1), 16 * n1 * n1, steps=n_c1, warmup=n_c1, best_of=1 if SHORT_SIDE else 3)
    # one call = one launch is bound by the host's submission rate (~3 us per step for a ~2 us kernel): the same loop captured ONCE into a
    # CUDA graph (cc_graph_begin / cc_graph_end) and replayed with one driver call per 50 evaluations shows the kernel itself
    try:
        per_graph = 50
        e1 = T.tanh(a1 * b1 + c1)
        e1.doBuffer().release()
        cuda.synchronize()
        with cuda.Graph() as g1:
            for _ in range(per_graph):
                e1.doBuffer().release()
        replays = 4 if SHORT_SIDE else 100
        ms, launches, _, _ = time_steps(cuda, g1.launch, replays, max(3, replays // 2))
        per = ms / (replays * per_graph)
        out["C1 tanh(a*b+c) 1024^2, 50 evaluations captured into one CUDA graph"] = {
            "ms": per, "kernels_per_step": launches / (replays * per_graph), "plan": 0, "gbs": 16 * n1 * n1 / per / 1e6,
            "frac_of_hbm": 16 * n1 * n1 / per / 1e6 / hbm_peak, "note": "16 MiB working set: L2-resident, so the HBM figure is only a yardstick"}
        g1.release()
        del e1
    except Exception as e:
        out["C1 tanh(a*b+c) 1024^2, 50 evaluations captured into one CUDA graph"] = {"error": str(e)[:200]}
    del a1, b1, c1
    x = T.random([ROWS, COLS], seed=5).doCache()
    measure("C3 full sum 16384^2", lambda: x.sum(), 4 * ROWS * COLS + 4)

    def axis(ax):
        parts = x.split(ax)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    # Tensor.sum of an inline expression folds its closure inside the reduction kernel (12 B/element, nothing materialised)
    a3, b3, c3 = (T.random([ROWS, COLS], seed=s).doCache() for s in (1, 2, 3))
    measure("sum of the C2 chain 16384^2, fused into one fold kernel", lambda: c2_chain(T, a3, b3, c3).sum(), 12 * ROWS * COLS + 4)
    del a3, b3, c3
    measure("C3 axis-0 sum 16384^2", lambda: axis(0), 4 * ROWS * COLS + 4 * COLS)
    measure("C3 axis-1 sum 16384^2", lambda: axis(1), 4 * ROWS * COLS + 4 * ROWS)
    del x
    n4 = 512
    t4 = T.random([n4, n4, n4], seed=7).doCache()
    m4 = T.random([n4, n4], seed=8).doCache()
    measure("C4 permute(2,0,1)+translate 512^3", lambda: t4.permute([2, 0, 1]).translate([3, -5, 7]), 8 * n4**3)
    measure("C4 trailing broadcast 512^2->512^3", lambda: m4.broadcast([n4, n4, n4]), 4 * n4**2 + 4 * n4**3)
    measure("C4 leading broadcast 512^2->512^3", lambda: m4.reshape([1, n4, n4]).broadcast([n4, n4, n4]), 4 * n4**2 + 4 * n4**3)
    measure("C4 split(1)/join round trip 512^3", lambda: T.join(t4.split(1)), 8 * n4**3)
    del t4, m4
    # the reference's convolution benchmark (benchmarks.scala:412-622) at its own size and at a realistic one
    for (cb, ch, cd) in ((128, 32, 8), (64, 56, 64)):
        try:
            ci, cw, cbias = (T.randomNormal(sh, seed=s).doCache() for sh, s in (([cb, ch, ch, cd], 1), ([3, 3, cd, cd], 2), ([cd], 3)))
            e = convolute(T, ci, cw, cbias)
            k = e.compile()
            kind = k.info.kind
            k.release()
            csteps = 10 if (SHORT_SIDE or cd > 8) else 500  # (the small one is a 7 us step: enough of them that the clock ramp is not what gets timed)
            ms, launches, _, _ = time_steps(cuda, lambda: e.doBuffer().release(), csteps, max(3, csteps // 10))
            per = ms / csteps
            flops = 2 * cb * ch * ch * cd * cd * 9
            out[f"convolution 3x3 batch {cb} {ch}x{ch} depth {cd} (benchmarks.scala:463-556)"] = {
                "ms": per, "tflops_fp32_equivalent": flops / per / 1e9, "gbs": 4 * 2 * cb * ch * ch * cd / per / 1e6, "plan": kind, "kernels_per_step": launches / csteps,
                "lowering": "warp-level 3xTF32 MMAs, weights in registers" if cd <= 8 else "implicit GEMM: gathered panels -> tcgen05 3xTF32"}
            del ci, cw, cbias, e
        except Exception as ex:
            out[f"convolution 3x3 batch {cb} {ch}x{ch} depth {cd} (benchmarks.scala:463-556)"] = {"error": str(ex)[:200]}
    n5 = 8192
    try:
        # C5 exactly as BASELINE.json words it: the matmul written as split / broadcast / sum (benchmarks.scala:188-191) through the lazy
        # Tensor API; the code generator re-rolls the 8192-term chain, sees the contraction through the fusion barrier and runs the
        # tcgen05 pipeline (the i*j*k product is never materialised)
        # dataset N: randomNormal(seed 9, 10); its singular pair (+inf, NaN — in the reference too) is zeroed on the host
        def dataset_n(seed):
            h = np.nan_to_num(T.randomNormal([n5, n5], seed=seed).flatArray(), nan=0.0, posinf=0.0, neginf=0.0).reshape(n5, n5)
            return T(h).doCache(), h

        (A, ha5), (B, hb5) = dataset_n(9), dataset_n(10)
        product = A.broadcast([n5, n5, n5]) * B.reshape([1, n5, n5]).broadcast([n5, n5, n5])
        parts = product.split(1)
        acc = parts[0]
        for p_ in parts[1:]:
            acc = acc + p_
        del parts
        k = acc.compile()
        kind = k.info.kind
        k.release()
        # value check BEFORE timing, on the kernel that is timed: sampled rows (both CTAs of a pair, first / middle / last tiles) against fp64
        # on the scale |A|.|B| (north star: <= 1e-5), and the whole result through the checksum identity sum(C) = colsum(A) . rowsum(B)
        c5 = acc.flatArray().reshape(n5, n5)
        rows5 = np.r_[0:2, 127:130, 255:257, 4095:4097, 8190:8192]
        a64, b64 = ha5[rows5].astype(np.float64), hb5.astype(np.float64)
        err5 = float((np.abs(c5[rows5].astype(np.float64) - a64 @ b64) / (np.abs(a64) @ np.abs(b64))).max())
        total, want_total = float(c5.astype(np.float64).sum()), float(ha5.astype(np.float64).sum(axis=0) @ b64.sum(axis=1))
        scale_total = float(np.abs(ha5).astype(np.float64).sum(axis=0) @ np.abs(b64).sum(axis=1))
        c5_check = {"max_err_over_absA_absB_sampled_rows": err5, "bar": 1e-5, "checksum_rel_err": abs(total - want_total) / scale_total,
                    "verified": bool(err5 <= 1e-5 and abs(total - want_total) <= 1e-5 * scale_total)}
        del c5, a64, b64, ha5, hb5

        def step():
            acc.doBuffer().release()

        # both operands fresh every step (hi/lo split of A and of B inside the timed region) ...
        cuda.set_operand_cache(False)
        ms, launches, _, _ = time_steps(cuda, step, 5, 2)
        per = ms / 5
        tf = 2 * n5**3 / per / 1e9
        out["C5 matmul 8192^3 as split/broadcast/sum (3xTF32 tcgen05, CTA pairs)"] = {
            "ms": per, "tflops": tf, "frac_of_3xtf32_peak": tf / tf_peak, "kernels_per_step": launches / 5, "plan": kind,
            "data": "dataset N: randomNormal(seed 9, 10), non-finite pair zeroed", "check": c5_check}
        # ... and with B unchanged between steps (weights): its panels are split once and kept by the runtime
        cuda.set_operand_cache(True)
        ms, launches, _, _ = time_steps(cuda, step, 5, 2)
        per = ms / 5
        tf = 2 * n5**3 / per / 1e9
        out["C5 matmul 8192^3, B unchanged between steps (panels cached)"] = {"ms": per, "tflops": tf, "frac_of_3xtf32_peak": tf / tf_peak,
                                                                             "kernels_per_step": launches / 5, "plan": kind}
        del acc, product
    except Exception as e:
        out["C5 matmul 8192^3 as split/broadcast/sum (3xTF32 tcgen05, CTA pairs)"] = {"error": str(e)[:200]}
    return out


def np_random(n: int, seed: int, first: int = 0) -> np.ndarray:
    """Tensor.random's stream on the host (Tensors.scala:106-117, 432-443): wang_hash(i ^ seed) / 2^32 for i in [first, first + n) — the
    checker of the sharded legs (numpy only: the cuda arm never touches oracle/)"""
    v = (np.arange(first, first + n, dtype=np.uint64) & 0xFFFFFFFF).astype(np.uint32) ^ np.uint32(seed & 0xFFFFFFFF)
    v = (v ^ np.uint32(61)) ^ (v >> np.uint32(16))
    v = v * np.uint32(9)
    v = v ^ (v << np.uint32(4))
    v = v * np.uint32(0x27D4EB2D)
    v = v ^ (v >> np.uint32(15))
    return (v.astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)


def np_dataset_e(n: int, seed: int, first: int = 0) -> np.ndarray:
    """dataset E: floor(random * 9) - 4, integers in {-4..5}: every summation order is exact"""
    return np.floor(np_random(n, seed, first) * np.float32(9.0)) - np.float32(4.0)


def sharded_configs(cuda, dist, rank: int, world: int, hbm_peak: float, tf_peak: float) -> dict:
    """C3 (16384^2, rows/N per GPU, partial sums combined over NVLink) and C5 (8192^3, A and C row-sharded, B replicated) at N GPUs, written
    as the SAME Tensor-API expressions as on one GPU over `.shard()`ed row blocks (include/compute_cuda.h: ct_shard / ct_gather). Strong
    scaling of the fixed BASELINE sizes; every time is the max over ranks of the device time including the exchange. Every result is
    compared with numpy on dataset E (exact in any order) BEFORE it is timed; `verified` records it."""
    import torch

    from compute.scala_b200 import sharding

    T = cuda.Tensor
    comm = sharding.Communicator(cuda, dist)
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())

    def all_sum_i64(a: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
        dist.all_reduce(t)
        return t.cpu().numpy()

    def all_true(ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def measure(name, expr, alg_bytes=None, flops=None, steps=10, verified=None):
        k = expr.compile() if hasattr(expr, "compile") else None

        def step():
            expr.doBuffer().release()

        for _ in range(3):
            step()
        cuda.synchronize()
        dist.barrier()
        cuda.timer_start()
        for _ in range(steps):
            step()
        ms = cuda.timer_stop() / steps
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rec = {"ms": ms, "n_gpus": world, "verified": verified, "distribution": expr.distribution}
        if flops:
            rec["tflops"] = flops / ms / 1e9
            rec["frac_of_3xtf32_peak_all_gpus"] = rec["tflops"] / (tf_peak * world)
        else:
            rec["gbs"] = alg_bytes / ms / 1e6
            rec["frac_of_hbm_all_gpus"] = rec["gbs"] / (hbm_peak * world)
        out[name] = rec
        return rec

    # ---- C3: dataset E, this rank's rows of the [16384, 16384] tensor --------------------------------------------------------------
    rows = sharding.shard_rows(ROWS, world, rank)[1]
    seed3 = 5 + 16 * rank
    r3 = T.random([rows, COLS], seed=seed3)
    nine, one, four = (T.fill(v, [rows, COLS]) for v in (9.0, 1.0, 4.0))
    x = ((r3 * nine) - (r3 * nine) % one - four).doCache().shard()
    hx = np_dataset_e(rows * COLS, seed3).reshape(rows, COLS).astype(np.int64)
    want_total = int(all_sum_i64(np.asarray([hx.sum()]))[0])
    want_cols = all_sum_i64(hx.sum(axis=0))
    want_rows = hx.sum(axis=1)
    total, col_sums, row_sums = x.sum(), comm.fold(x.split(0)), comm.fold(x.split(1))
    row_sums_gathered = row_sums.gather()
    had_peer = comm.peer
    for route, tag in ((True, "fused / one-shot over NVLink peer memory"), (False, "NCCL")):
        if route and not comm.peer:
            continue
        comm.route_peer(route)
        ok = all_true(float(total.flatArray()[0]) == float(want_total))
        measure(f"C3 full sum 16384^2 sharded + allreduce(1 float) [{tag}]", total, alg_bytes=4 * ROWS * COLS, steps=50, verified=ok)
        ok = all_true(np.array_equal(col_sums.flatArray().astype(np.int64), want_cols))
        measure(f"C3 axis-0 sum 16384^2 sharded + allreduce(16384 floats) [{tag}]", col_sums, alg_bytes=4 * ROWS * COLS, steps=50, verified=ok)
        got = row_sums_gathered.flatArray().astype(np.int64)
        ok = all_true(np.array_equal(got[rank * rows:(rank + 1) * rows], want_rows) and int(got.sum()) == want_total) if ROWS % world == 0 else None
        measure(f"C3 axis-1 sum 16384^2 sharded + allgather [{tag}]", row_sums_gathered, alg_bytes=4 * ROWS * COLS, steps=50, verified=ok)
    if had_peer:
        comm.route_peer(True)
    del x, total, col_sums, row_sums, row_sums_gathered, r3, nine, one, four, hx
    # ---- C5: A and C row-sharded, B replicated; the reference's formulation (benchmarks.scala:188-191) on this rank's rows ---------------
    n5 = 8192
    m5 = sharding.shard_rows(n5, world, rank)[1]
    if m5 > 0 and n5 % world == 0:
        def e_tensor(shape, seed):
            r = T.random(shape, seed=seed)
            return ((r * T.fill(9.0, shape)) - (r * T.fill(9.0, shape)) % T.fill(1.0, shape) - T.fill(4.0, shape)).doCache()

        seed_a = lambda r: 9 + 16 * r  # noqa: E731
        A = e_tensor([m5, n5], seed_a(rank)).shard()
        B = e_tensor([n5, n5], 10)
        hb = np_dataset_e(n5 * n5, 10).reshape(n5, n5).astype(np.float64)
        c = comm.matmul_pattern(A, B)
        kinfo = c.compile().info

        def rows_ok(got_rows: np.ndarray, owner: int, local_rows) -> bool:
            ha = np.stack([np_dataset_e(n5, seed_a(owner), first=int(r) * n5) for r in local_rows]).astype(np.float64)
            return bool(np.array_equal(got_rows.astype(np.float64), ha @ hb))

        sample = np.r_[0:2, 127:129, m5 - 2:m5]
        got = c.flatArray().reshape(m5, n5)
        ok = all_true(rows_ok(got[sample], rank, sample))
        rec = measure("C5 matmul 8192^3 as split/broadcast/sum on row blocks (B replicated, C left sharded)", c, flops=2 * n5**3, steps=5, verified=ok)
        rec["plan"] = int(kinfo.kind)
        del got
        for route, tag in ((True, "fused into the contraction's epilogue: TMA stores over NVLink peer memory"), (False, "contraction, then ncclAllGather")):
            if route and not comm.peer:
                continue
            comm.route_peer(route)
            cg = c.gather(zero_copy=True)
            whole = cg.flatArray().reshape(n5, n5)
            ok = all_true(all(rows_ok(whole[o * m5 + sample], o, sample) for o in range(world)))
            del whole
            rec = measure(f"C5 matmul 8192^3 row-sharded + allgather(C) [{tag}]", cg, flops=2 * n5**3, steps=5, verified=ok)
            rec["plan"] = int(kinfo.kind)
            del cg
        if had_peer:
            comm.route_peer(True)
        del A, B, c
    cuda.synchronize()
    comm.close()
    return out


def run_cuda(args, rank: int, local_rank: int, world: int):
    from compute.scala_b200 import cuda

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    cuda.init(local_rank, streams=1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    T = cuda.Tensor
    pk, pk_kind = peaks()
    hbm_peak = float(pk["hbm_gbs"])
    rows = args.rows
    n = rows * COLS
    shape = [rows, COLS]
    alg_bytes = BYTES_PER_ELEMENT * n

    # weak scaling: rank r owns rows [r*rows, (r+1)*rows) of a [world*rows, 16384] tensor; elementwise graphs need no exchange.
    # random()'s stream is a function of the global element index, so shard r is seeded to continue it: i ^ seed with the
    # global offset folded into distinct seeds keeps shards independent (synthetic data either way).
    seeds = [1 + 16 * rank, 2 + 16 * rank, 3 + 16 * rank]
    a, b, c = (T.random(shape, seed=s).doCache() for s in seeds)
    expr = c2_chain(T, a, b, c)
    kern = expr.compile()
    kinfo = kern.info
    assert kinfo.algorithmic_bytes == alg_bytes, (kinfo.algorithmic_bytes, alg_bytes)

    def step():
        expr.doBuffer().release()

    if dist:
        dist.barrier()
    ms, launches, w0, w1 = time_steps(cuda, step, args.steps, args.warmup)
    if dist:
        import torch

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    ms_per_step = ms / args.steps
    value = world * alg_bytes / ms_per_step / 1e6  # GB/s, whole job
    kernel_gbs = alg_bytes / ms_per_step / 1e6

    # ---- e2e: host buffers in, host buffer out, through the same public API -------------------------------------------------
    chunks = args.e2e_chunks
    crow = rows // chunks
    cn = crow * COLS
    ha, hb, hc, ho = (cuda.PinnedArray(n) for _ in range(4))
    for h, src in ((ha, a), (hb, b), (hc, c)):
        src.flatArrayInto(h.ptr, n)
    da, db, dc = ([cuda.Buffer.alloc(cn) for _ in range(chunks)] for _ in range(3))
    exprs = []
    for i in range(chunks):
        ta, tb, tc = (T.fromBuffer(d[i], [crow, COLS]) for d in (da, db, dc))
        exprs.append(c2_chain(T, ta, tb, tc))

    def e2e_step():
        # leading-axis chunks: H2D of chunk i+1 overlaps the kernel of chunk i and the D2H of chunk i-1 (separate streams, event-ordered)
        for i in range(chunks):
            off = i * cn * 4
            da[i].upload(ha.ptr + off, cn)
            db[i].upload(hb.ptr + off, cn)
            dc[i].upload(hc.ptr + off, cn)
            out = exprs[i].doBuffer()
            out.to_host_async(ho.ptr + off, cn)
            out.release()
        cuda.synchronize()  # the result is in host memory when the step ends

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    if dist:
        dist.barrier()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    if dist:
        import torch

        t = torch.tensor([e2e_sec], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_value = world * alg_bytes / e2e_sec / 1e9
    e2e_ok = None
    if rank == 0:
        # the bytes that came back are the chain's result (spot check against the device-resident evaluation)
        ref_out = expr.flatArray()
        e2e_ok = bool(np.array_equal(ref_out[: 1 << 20].view(np.uint32), ho.array[: 1 << 20].view(np.uint32)))
    # The ceiling of that number on this box: the same bytes (3 inputs up, 1 result down, both directions at once, every rank at once) with
    # no kernel at all. It separates what the host / PCIe fabric gives N processes from what the staging design costs.
    def copies_only():
        for i in range(chunks):
            off = i * cn * 4
            da[i].upload(ha.ptr + off, cn)
            db[i].upload(hb.ptr + off, cn)
            dc[i].upload(hc.ptr + off, cn)
            da[i].to_host_async(ho.ptr + off, cn)
        cuda.synchronize()

    copies_only()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        copies_only()
    copy_sec = (time.perf_counter() - t0) / 3
    copy_sec_max = copy_sec
    if dist:
        import torch

        t = torch.tensor([copy_sec], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_sec_max = float(t.item())
    e2e_ceiling = {"value": world * alg_bytes / copy_sec_max / 1e9, "unit": "GB/s", "h2d_gbs_per_rank": 3 * n * 4 / copy_sec_max / 1e9,
                   "d2h_gbs_per_rank": n * 4 / copy_sec_max / 1e9,
                   "what": "the step's copies alone (3 GiB up + 1 GiB down per rank, all ranks at once, no kernel): the host / PCIe ceiling of e2e on this box"}
    time.sleep(0.1)
    clocks = sampler.summary(w0, w1)
    sampler.stop()

    line = None
    if rank == 0:
        tf_peak = float(pk.get("bf16_tflops", 1590.0)) / 2 / 3  # 3xTF32 fp32-equivalent peak = TF32 peak / 3 ~= bf16 / 2 / 3
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C2_WORKLOAD, "shape_per_gpu": shape, "elements_per_gpu": n,
                       "inputs": "Tensor.random seeds 1,2,3 (Wang hash), cached in HBM", "l2": "inputs (3 GiB) and output (1 GiB) are far larger than the 126 MB L2",
                       "sharding": "leading axis, no collective", "e2e_chunks": chunks},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": kernel_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": kernel_gbs / hbm_peak,
                         "traffic": None, "peak_source": pk_kind, "kernel": "jit_kernel (fused elementwise template)",
                         "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": 3 * n * 4, "d2h_bytes_per_step": n * 4, "ms_per_step": e2e_sec * 1e3,
                    "steps": e2e_steps, "result_matches_device_path": e2e_ok, "h2d_gbs_per_rank": 3 * n * 4 / e2e_sec / 1e9,
                    "copy_ceiling": e2e_ceiling},
        }
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel cannot be measured by a plain run: the figure is the constant that the
        # committed `ncu --set full` capture of the same kernel / same shape reports (profiles/), labelled as such
        traffic_file = os.path.join(ROOT, "profiles", "c2_traffic_bytes.json")
        if os.path.exists(traffic_file) and rows == ROWS:
            try:
                tj = json.load(open(traffic_file))
                line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = "constant from profiles/c2_traffic_bytes.json <- " + str(tj.get("source", "ncu --set full capture")) + " (not measured in this run)"
            except Exception:
                pass
    for h in (ha, hb, hc, ho):
        h.free()
    del exprs, da, db, dc
    if rank == 0 and world == 1:
        if not args.no_side_configs:
            del a, b, c, expr
            line["configs"] = side_configs(cuda, hbm_peak, tf_peak)
        if not args.no_cpu_baseline:  # last: ~20 s during which the GPU idles and drops its clocks
            log2n = int(os.environ.get("BENCH_REFERENCE_SAMPLE_LOG2", "28"))
            times, cores, _ = cpu_c2(1 << log2n, 1, 40, budget_s=10.0)
            sec = sum(times) / len(times)
            line["cpu_baseline"] = {"value": BYTES_PER_ELEMENT * (1 << log2n) / sec / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                                    "sample": f"all 2^{log2n} elements per pass (the cuda arm's workload), mean of {len(times)} passes "
                                              f"({sum(times):.1f} s of wall clock on {cores} threads); " + CPU_NOTE}
            if not args.no_side_configs and log2n == 28:
                try:
                    for name, rec in cpu_side_baselines().items():
                        for key in line.get("configs", {}):
                            if key.startswith(name):
                                line["configs"][key]["cpu_baseline"] = rec
                except Exception as e:
                    line["cpu_baseline"]["side_error"] = str(e)[:200]
    if world > 1 and not args.no_side_configs:
        pk2, _ = peaks()
        sc = sharded_configs(cuda, dist, rank, world, float(pk2["hbm_gbs"]), float(pk2.get("bf16_tflops", 1590.0)) / 6)
        if rank == 0:
            line["configs"] = sc
    if rank == 0:
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """the ONE JSON line of the contract, on the process' real stdout"""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries underneath (NCCL prints its version banner on stdout when the box sets NCCL_DEBUG) must not add lines to stdout:
    # everything written to fd 1 from here on goes to stderr, and only emit() writes to the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS, help="rows of the [rows,16384] shard per GPU (default = the full config)")
    ap.add_argument("--e2e-chunks", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true")
    ap.add_argument("--short-side", action="store_true", help="few steps per side config (for ncu launch lists)")
    args = ap.parse_args()
    if args.warmup < 3:
        sys.stderr.write(f"bench.py: --warmup {args.warmup} raised to 3 (the timing rules require W >= 3); the JSON line reports 3\n")
        args.warmup = 3
    global SHORT_SIDE
    SHORT_SIDE = args.short_side
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_cuda(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
