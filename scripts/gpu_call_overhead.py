"""Where the ~25 us of a tiny `flatBuffer` call go (host side): graph-to-launch, launch + wait, read-back."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from compute.scala_b200 import cuda  # noqa: E402

cuda.init(0)
T = cuda.Tensor
a, b, c = (T.random([32, 32], seed=s).doCache() for s in (1, 2, 3))
e = T.tanh(a * b + c)
e.flatArray()
N = 2000


def timed(name, fn):
    for _ in range(200):
        fn()
    cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N):
        fn()
    cuda.synchronize()
    print(f"{name:48s} {(time.perf_counter() - t0) / N * 1e6:7.2f} us/call", flush=True)


timed("python no-op ctypes call (cc_version)", lambda: cuda._L().cc_version())
timed("doBuffer + release (launch only, async)", lambda: e.doBuffer().release())


def launch_sync():
    e.doBuffer().release()
    cuda.synchronize()


timed("doBuffer + release + synchronize", launch_sync)
buf = e.doBuffer()
pin = cuda.PinnedArray(1024)
timed("to_host_async of a ready buffer + synchronize", lambda: (buf.to_host_async(pin.ptr, 1024), cuda.synchronize()))
timed("flatArrayInto pinned", lambda: e.flatArrayInto(pin.ptr, 1024))
timed("flatBuffer + release", lambda: e.flatBuffer().release())
timed("flatArray", lambda: e.flatArray())
s = a.sum()
timed("sum().flatBuffer", lambda: a.sum().flatBuffer().release())
print(cuda.stats())
