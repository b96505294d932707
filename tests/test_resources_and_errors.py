"""GPU: resource limits and error behaviour at the boundary (OpenCL.scala:143-312 error codes -> typed exceptions; deterministic
release, README.md:10; monadicClose, OpenCL.scala:1331-1337).  The shutdown / re-init cycle runs in its own process."""
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    from compute.scala_b200 import cuda as c

    c.init()
    yield c
    c.synchronize()


def test_out_of_memory_is_an_error_not_a_crash(cuda):
    T = cuda.Tensor
    with pytest.raises(cuda.ComputeCudaError) as e:
        cuda.Buffer.alloc(60 * 1000**3)  # 240 GB > 180 GB of HBM
    assert e.value.status == -9  # CC_ERR_OUT_OF_MEMORY (OutOfResources / MemObjectAllocationFailure, O:160-170)
    # an expression whose OUTPUT cannot be allocated: same error from the slow action, operands released
    base = cuda.stats()["bytes_in_use"]
    big = T.random([1 << 16, 1 << 15], seed=1).broadcast([1 << 16, 1 << 15, 64]) * T.fill(2.0, [1 << 16, 1 << 15, 64])  # 2^37 floats
    with pytest.raises(cuda.ComputeCudaError) as e:
        big.doBuffer()
    assert e.value.status in (-9, -1)
    del big
    cuda.synchronize()
    assert cuda.stats()["bytes_in_use"] == base
    # the library is still usable
    assert T.fill(3.0, [4]).sum().flatArray().tolist() == [12.0]


def test_the_pool_gives_memory_back_under_pressure(cuda):
    """freed buffers stay pooled; an allocation that does not fit trims the pool instead of failing"""
    info = cuda.device_info()
    total = int(info.total_mem)
    chunk = (total // 3) // 4  # floats: a third of the device each
    a = cuda.Buffer.alloc(chunk)
    b = cuda.Buffer.alloc(chunk)
    a.release()
    b.release()
    assert cuda.stats()["bytes_pooled"] >= 2 * chunk * 4 - (8 << 20)
    c = cuda.Buffer.alloc(chunk * 2 + (1 << 20))  # needs the pooled two thirds back
    assert c.length == chunk * 2 + (1 << 20)
    c.release()
    big = cuda.Buffer.alloc(chunk * 2 + (1 << 20))  # pool hit this time
    big.release()
    cuda.memory_trim()  # do not sit on two thirds of the device for the rest of the session
    assert cuda.stats()["bytes_pooled"] == 0


def test_invalid_handles_and_arguments(cuda):
    import ctypes as C

    L = cuda._L()
    u64 = C.c_uint64
    assert L.cc_buffer_release(u64(0xDEADBEEF)) == -1
    assert L.cc_event_wait(u64(12345)) == -1
    assert L.cc_kernel_release(u64(1)) == -1
    assert L.ct_release(u64(0)) == -1 and L.ct_release(u64(0x1000)) == -1  # null / made-up tensor handles: IllegalArgument, no crash
    t = cuda.Tensor.fill(1.0, [2])
    h = t._h
    t.release()
    assert L.ct_release(u64(int(h.value) if hasattr(h, "value") else int(h))) == -1  # released twice
    out = u64()
    assert L.cc_buffer_alloc(u64(16), None) == -1
    assert L.cc_compile(b"\x00" * 8, u64(8), C.byref(out)) == -6  # CC_ERR_BAD_TREE
    b = cuda.Buffer.alloc(16)
    with pytest.raises(cuda.ComputeCudaError):
        b.to_host(17)  # read past the end
    k = cuda.Tensor.tanh(cuda.Tensor.random([16], seed=1)).compile()
    small = cuda.Buffer.alloc(8)
    ev = u64()
    args = (u64 * 1)(b.handle)
    st = L.cc_launch(u64(k.handle), args, 1, u64(small.handle), None, 0, C.byref(ev))
    assert st == -1 and b"output buffer" in L.cc_last_error()  # output too small: refused before any launch
    st = L.cc_launch(u64(k.handle), args, 1, u64(b.handle), None, 0, C.byref(ev))
    assert st == -1 and b"aliases" in L.cc_last_error()
    b.release(), small.release()


def test_shutdown_and_reinit_in_a_fresh_process():
    code = textwrap.dedent(
        """
        import numpy as np
        from compute.scala_b200 import cuda
        T = cuda.Tensor
        cuda.init(0)
        x = T.random([64, 64], seed=1).doCache()
        want = T.tanh(x * x).flatArray()
        k = T.tanh(x * x).compile()
        held = x.doBuffer()
        cuda.shutdown()                      # monadicClose: cache, pools, streams, context (O:1331-1337)
        assert not cuda.is_initialized()
        try:
            T.fill(1.0, [4]).flatArray()
            raise SystemExit("evaluated without a context")
        except cuda.ComputeCudaError as e:
            assert e.status == -2
        cuda.init(0)                         # a second life: new context, empty caches
        # (the refused evaluation above still compiled its kernel: NVRTC needs no context, which is what the CPU-only tests rely on)
        assert cuda.kernel_cache_size() == 1, cuda.kernel_cache_size()
        assert cuda.stats()["bytes_in_use"] <= (1 << 20), cuda.stats()  # nothing but the runtime's own scratch
        y = T.random([64, 64], seed=1)
        got = T.tanh(y * y).flatArray()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        try:
            held.to_host(4)                  # a handle of the previous life: its memory went with the context
            raise SystemExit("stale buffer was readable")
        except cuda.ComputeCudaError:
            pass
        held.release(); del x, k
        cuda.shutdown(); cuda.shutdown()     # idempotent
        print("ok")
        """
    )
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-2000:], r.stderr[-2000:])


def test_ping_pong_chains_over_recycled_pool_blocks_stay_exact(cuda):
    """tiny producer -> consumer chains whose buffers ping-pong between two or three recycled pool blocks, with programmatic
    dependent launch on (the default). Every step adds exactly 1, so a kernel that ever read a stale or half-written input (a non-coherent
    load of data its early-resident grid overlapped with) would leave the final value short. Streaming loads, L1-allocating loads of a
    broadcast operand, and a two-stream variant."""
    T = cuda.Tensor
    steps = 3000
    for n in (64, 4096, 65536):
        x = T.fill(0.0, [n]).doCache()
        one = T.fill(1.0, [n])
        for _ in range(steps):
            x = (x + one).doCache()
        got = x.flatArray()
        assert (got == float(steps)).all(), (n, got[:8])
    rows, cols = 64, 256
    x = T.fill(0.0, [rows, cols]).doCache()
    row = T(np.ones(rows, np.float32))  # broadcast along the columns: read through the L1-allocating path
    for _ in range(steps):
        x = (x + row.broadcast([rows, cols])).doCache()
    assert (x.flatArray() == float(steps)).all()
    # reductions in the chain: the scalar result feeds the next step
    s = T.fill(1.0, [1024]).doCache()
    acc = T.scalar(0.0).doCache()
    for _ in range(300):
        acc = (acc + s.sum()).doCache()
    assert acc.flatArray()[0] == 300 * 1024.0
