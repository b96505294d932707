"""TEST INFRASTRUCTURE ONLY (see README.md).  Child process of tests/test_driver_calls.py: runs one scenario through the public API with
the driver spy in front of libcuda (LD_LIBRARY_PATH) and prints the spy's counters as one JSON line.  Kernels never execute here."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from compute.scala_b200 import cuda  # noqa: E402

SPY = ctypes.CDLL("libcuda.so.1")
assert hasattr(SPY, "spy_report"), "the real driver is in front of the spy"


def report() -> dict:
    buf = ctypes.create_string_buffer(1 << 14)
    SPY.spy_report(buf, 1 << 14)
    return json.loads(buf.value)


def reset() -> None:
    SPY.spy_reset()


T = cuda.Tensor


def leaf(shape, value=1.0):
    return T(np.full(shape, value, np.float32)).doCache()


def chain(parts):
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def steady_loop():
    a, b, c = (leaf([64, 64]) for _ in range(3))
    e = T.tanh(a * b + c)
    for _ in range(8):
        e.doBuffer().release()
    reset()
    s0 = cuda.stats()
    for _ in range(100):
        e.doBuffer().release()
    s1 = cuda.stats()
    return {**report(), "alloc_calls": s1["alloc_calls"] - s0["alloc_calls"], "pool_hits": s1["pool_hits"] - s0["pool_hits"], "compiles": s1["compiles"] - s0["compiles"]}


def independent_rotate():
    a = leaf([64, 64])
    exprs = [T.abs(a) * T.fill(float(i + 2), [64, 64]) for i in range(8)]
    for e in exprs:
        e.compile().release()
    cuda.synchronize()
    reset()
    held = [e.doBuffer() for e in exprs]  # fresh outputs, nothing in common but a long-synced input: the commands rotate over the streams
    r = report()
    for h in held:
        h.release()
    return r


def first_use_of_uploaded_inputs():
    a, b, c = (leaf([64, 64]) for _ in range(3))  # three H2D copies on the copy stream
    e = a * b + c
    e.compile().release()
    reset()
    first = e.doBuffer()
    r1 = report()
    reset()
    first.release()
    e.doBuffer().release()
    r2 = report()
    return {"first": r1, "second": r2}


def read_back():
    a = leaf([32, 32])
    small = T.abs(a) + a
    big_in = leaf([256, 256])
    big = T.abs(big_in) + big_in
    small.flatBuffer().release()
    big.flatBuffer().release()
    reset()
    small.flatBuffer().release()
    rs = report()
    reset()
    big.flatBuffer().release()
    rb = report()
    return {"small": rs, "big": rb}


def read_back_values():
    """the spy's device-to-host copy writes 42.0f into every float it is asked to deliver"""
    x = leaf([300, 256])
    e = T.abs(x) + x
    arr = e.flatArray()
    hb = e.flatBuffer()
    in_scope = hb.array
    out = {"flat_array": [str(arr.dtype), list(arr.shape), float(arr.min()), float(arr.max())],
           "flat_buffer": [str(in_scope.dtype), list(in_scope.shape), float(in_scope.min()), float(in_scope.max()), hb.n]}
    hb.release()
    out["released"] = hb.array is None
    hb.release()  # idempotent
    pin = cuda.PinnedArray(300 * 256 + 8)
    pin.array[:] = -1.0
    e.flatArrayInto(pin.ptr, 300 * 256)
    out["into"] = [float(pin.array[: 300 * 256].min()), float(pin.array[: 300 * 256].max()), float(pin.array[300 * 256:].max())]
    pin.free()
    small = (leaf([4, 4]) * leaf([4, 4])).flatArray()  # stored by the kernel itself (which never runs here): shape only
    out["small"] = [str(small.dtype), list(small.shape)]
    empty = leaf([0, 4])
    out["empty"] = [list((T.abs(empty)).flatArray().shape), T.abs(empty).flatBuffer().n]
    with pytest_raises_illegal():
        T.fill(1.0, [2, 3]) + T.fill(1.0, [4, 5])
    out["typed_error"] = True
    out["native_binding"] = bool(cuda._hot().native)
    return out


class pytest_raises_illegal:
    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        assert et is cuda.IllegalArgumentException, et
        return True


def two_launch_plan_and_fold():
    x = leaf([300, 64])
    e = chain(x.split(0))
    f = (x * x).sum()
    for _ in range(4):
        e.doBuffer().release()
        f.doBuffer().release()
    reset()
    for _ in range(50):
        e.doBuffer().release()
    re_ = report()
    reset()
    for _ in range(50):
        f.doBuffer().release()
    rf = report()
    return {"axis": re_, "fold": rf, "axis_launches_per_step": e.compile().info.n_launches}


def graph_capture_and_replay():
    """cc_graph_*: a loop's worth of evaluations recorded once (kernel launches go to the capture stream, nothing else does), then one
    cuGraphLaunch per replay; memory the captured commands used never returns to the pool while the graph lives"""
    a, b, c = leaf([64, 64]), leaf([64, 64], 2.0), leaf([64, 64], 3.0)
    e = cuda.Tensor.tanh(a * b + c)
    x = leaf([300, 64])
    col = chain(x.split(0))
    e.doBuffer().release()
    col.doBuffer().release()
    cuda.synchronize()
    reset()
    s0 = cuda.stats()
    with cuda.Graph() as g:
        for _ in range(20):
            e.doBuffer().release()  # the same pooled block every time: the capture's own pool
        kept = col.doBuffer()  # a result the caller keeps: refreshed by every replay
        refused = False
        try:
            kept.to_host(64)  # a copy cannot be part of the graph
        except cuda.ComputeCudaError as err:
            refused = "cannot be captured" in str(err)
    captured = report()
    s1 = cuda.stats()
    reset()
    for _ in range(5):
        g.launch()
    replay = report()
    s2 = cuda.stats()
    in_use_with_graph = s2["bytes_in_use"]
    commands = g.commands
    g.release()
    kept.release()
    e.doBuffer().release()  # the runtime is back to normal launches
    s3 = cuda.stats()
    return {"captured": captured, "replay": replay, "commands": commands, "refused_copy": refused,
            "kernels_counted_while_capturing": s1["device_kernels"] - s0["device_kernels"],
            "kernels_counted_by_replays": s2["device_kernels"] - s1["device_kernels"], "launch_after": s3["launches"] - s2["launches"]}


def pdl_only_when_inputs_are_settled():
    """an early-resident (PDL) grid may not read, through non-coherent loads, what the command right before it is still writing: a kernel that
    consumes the previous kernel's output is launched in plain stream order; loops over long-lived inputs keep the attribute"""
    a, b = leaf([64, 64]), leaf([64, 64], 2.0)
    e = a * b + a
    e.doBuffer().release()
    reset()
    for _ in range(20):
        e.doBuffer().release()  # same old inputs every step
    settled = report()
    x = a
    reset()
    for _ in range(20):
        x = (x * b).doCache()  # each step reads the previous step's output
    chained = report()
    return {"settled": settled, "chained": chained}


def structural_cache():
    a, b = leaf([16, 16]), leaf([16, 16], 2.0)
    reset()
    s0 = cuda.stats()
    (a + b).doBuffer().release()
    (b + a).doBuffer().release()  # same structure, other parameters (alpha-equivalent: Trees.scala:23-177)
    (a * b).doBuffer().release()
    s1 = cuda.stats()
    return {**report(), "compiles": s1["compiles"] - s0["compiles"], "cache_hits": s1["cache_hits"] - s0["cache_hits"]}


def balance_on_shutdown():
    a, b, c = (leaf([64, 64]) for _ in range(3))
    e = T.tanh(a * b + c)
    for _ in range(20):
        e.doBuffer().release()
    e.flatArray()
    chain(a.split(0)).flatArray()
    big = leaf([300, 64])
    chain(big.split(0)).flatArray()  # a split axis reduction (second stage: its own launch, or fused with CC_FUSE_COL_STAGE=1)
    big.release()
    (a * b).sum().flatBuffer().release()
    del e
    a.release(), b.release(), c.release()
    live = cuda.live_tensors()
    in_use = cuda.stats()["bytes_in_use"]
    cuda.shutdown()
    return {**report(), "live_tensors": live, "bytes_in_use_before_shutdown": in_use}


def threads():
    """calls come from arbitrary threads (ExecutionContext.global, OpenCL.scala:414-416; JMH Threads.MAX, benchmarks.scala:56): eight
    threads build, evaluate, read back and release concurrently (the native binding releases the GIL around the calls)"""
    import threading

    a, b = leaf([64, 64]), leaf([64, 64], 2.0)
    shared = T.tanh(a * b)
    shared.doBuffer().release()
    cuda.synchronize()
    base = cuda.stats()
    reset()
    failures = []

    def work(tid):
        try:
            for i in range(200):
                shared.doBuffer().release()                       # one plan, many threads
                e = T.abs(a + T.fill(float(tid), [64, 64])) * b   # per-thread structure (the literal is part of the key)
                if i % 4 == 0:
                    assert e.flatArray().shape == (4096,)
                elif i % 4 == 1:
                    e.flatBuffer().release()
                else:
                    e.doBuffer().release()
                e.release()
        except Exception as ex:  # noqa: BLE001
            failures.append(repr(ex)[:300])

    ts = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    hung = sum(t.is_alive() for t in ts)
    cuda.synchronize()
    s1 = cuda.stats()
    return {**report(), "failures": failures, "hung": hung, "launches": s1["launches"] - base["launches"], "compiles": s1["compiles"] - base["compiles"],
            "bytes_in_use_delta": s1["bytes_in_use"] - base["bytes_in_use"], "live_tensors": cuda.live_tensors()}


def compile_does_not_block_launches():
    """one thread JIT-compiles new structures (tens of ms each, outside the runtime lock) while another keeps stepping a cached plan"""
    import threading
    import time

    a, b = leaf([64, 64]), leaf([64, 64], 2.0)
    e = T.tanh(a * b)
    e.doBuffer().release()
    stop, worst, steps, compile_ms = [False], [0.0], [0], []

    def stepper():
        while not stop[0]:
            t0 = time.perf_counter()
            e.doBuffer().release()
            worst[0] = max(worst[0], time.perf_counter() - t0)
            steps[0] += 1

    th = threading.Thread(target=stepper)
    th.start()
    time.sleep(0.02)
    for i in range(6):
        big = a
        for j in range(12):
            big = T.tanh(big * b + T.fill(float(i * 100 + j), [64, 64]))
        t0 = time.perf_counter()
        big.compile().release()
        compile_ms.append((time.perf_counter() - t0) * 1e3)
    stop[0] = True
    th.join()
    return {"worst_step_ms": worst[0] * 1e3, "steps": steps[0], "compile_ms": compile_ms}


def same_structure_from_many_threads():
    import threading

    a = leaf([64, 64])
    s0 = cuda.stats()
    e = T.exp(a) * T.fill(3.0, [64, 64]) - a
    failures = []

    def work():
        try:
            e.doBuffer().release()
            (T.exp(a) * T.fill(3.0, [64, 64]) - a).doBuffer().release()  # a structurally equal twin built by this thread
        except Exception as ex:  # noqa: BLE001
            failures.append(repr(ex)[:200])

    ts = [threading.Thread(target=work) for _ in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    s1 = cuda.stats()
    return {**report(), "failures": failures, "compiles": s1["compiles"] - s0["compiles"], "nvrtc_compiles": s1["nvrtc_compiles"] - s0["nvrtc_compiles"],
            "launches": s1["launches"] - s0["launches"]}


def arm(name: str, skip: int, code: int, times: int = 1) -> None:
    SPY.spy_arm.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_int, ctypes.c_long]
    SPY.spy_arm(name.encode(), skip, code, times)


def faults():
    """one driver entry point is armed to fail in the middle of the work; the scenario keeps going and reports what the caller saw"""
    case = sys.argv[2]
    errors, ok = [], 0
    a, b, c = (leaf([64, 64]) for _ in range(3))
    e = T.tanh(a * b + c)
    e.doBuffer().release()
    cuda.synchronize()
    base = cuda.stats()["bytes_in_use"]

    def attempt(fn):
        nonlocal ok
        try:
            fn()
            ok += 1
        except cuda.ComputeCudaError as ex:
            errors.append([ex.status, str(ex)[:160]])

    if case == "launch":
        arm("cuLaunchKernelEx", 12, 719)
        for _ in range(30):
            attempt(lambda: e.doBuffer().release())
    elif case == "module":
        arm("cuModuleLoadData", 0, 200)
        f = T.exp(a) - b
        for _ in range(5):
            attempt(lambda: f.doBuffer().release())  # the failed load is retried by the next evaluation
        del f
    elif case in ("alloc_once", "alloc_always"):
        arm("cuMemAlloc", 2, 2, 1 if case == "alloc_once" else 1000)  # CUDA_ERROR_OUT_OF_MEMORY
        def fresh(i):  # new size classes: the input and the result each need a fresh cuMemAlloc
            x = leaf([128 << i, 32])
            try:
                (T.abs(x) + x).doBuffer().release()
            finally:
                x.release()

        for i in range(6):
            attempt(lambda: fresh(i))
        arm("cuMemAlloc", 0, 2, 0)  # disarm
    elif case in ("d2h", "sync"):
        big = leaf([300, 256])
        g = T.abs(big) + big
        g.flatArray()
        arm("cuMemcpyDtoHAsync" if case == "d2h" else "cuEventSynchronize", 1, 719)
        for _ in range(4):
            attempt(lambda: g.flatArray())
        for _ in range(4):
            attempt(lambda: g.flatBuffer().release())
        big.release()
        del g
    cuda.synchronize()
    mid = cuda.stats()["bytes_in_use"]
    after_ok = 0
    for _ in range(5):  # the runtime is still usable after the failure
        try:
            e.doBuffer().release()
            after_ok += 1
        except cuda.ComputeCudaError:
            pass
    del e
    a.release(), b.release(), c.release()
    out = {"errors": errors, "ok": ok, "after_ok": after_ok, "bytes_in_use_base": base, "bytes_in_use_after": mid, "live_tensors": cuda.live_tensors()}
    cuda.shutdown()
    out.update(report())
    return out


if __name__ == "__main__":
    cuda.init(0)
    out = globals()[sys.argv[1]]()
    print(json.dumps(out))
