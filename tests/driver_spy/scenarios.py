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
    (a * b).sum().flatBuffer().release()
    del e
    a.release(), b.release(), c.release()
    live = cuda.live_tensors()
    in_use = cuda.stats()["bytes_in_use"]
    cuda.shutdown()
    return {**report(), "live_tensors": live, "bytes_in_use_before_shutdown": in_use}


if __name__ == "__main__":
    cuda.init(0)
    out = globals()[sys.argv[1]]()
    print(json.dumps(out))
