"""GPU: the boundary's threading contract. The reference is driven by arbitrary threads (ExecutionContext.global,
OpenCL.scala:414-416; JMH Threads.MAX, benchmarks.scala:56), orders commands by event wait lists rather than submission order
(Tensors.scala:1363,1374) and resumes continuations from driver threads (clSetEventCallback, OpenCL.scala:379-392, 1246-1263)."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    from compute.scala_b200 import cuda as c

    c.init()
    yield c
    c.synchronize()


def test_concurrent_callers(cuda):
    T = cuda.Tensor
    errors = []

    def worker(tid):
        try:
            rng = np.random.default_rng(tid)
            for it in range(25):
                n = 48 + 8 * tid + it % 3
                a = rng.integers(-4, 5, (n, n)).astype(np.float32)
                b = rng.integers(-4, 5, (n, n)).astype(np.float32)
                ta, tb = T(a), T(b)
                got = (ta * tb + T.fill(float(tid), [n, n])).flatArray().reshape(n, n)
                assert np.array_equal(got, a * b + np.float32(tid)), ("fused", tid, it)
                assert (ta + tb).sum().flatArray()[0] == np.float32((a + b).astype(np.int64).sum()), ("sum", tid, it)
                # out[i, j] = a[j + 1, i - 1] or padding 0 (permute then translate, one composed matrix)
                v = ta.permute([1, 0]).translate([1, -1]).flatArray().reshape(n, n)
                want = np.zeros((n, n), np.float32)
                want[1:, : n - 1] = a.T[: n - 1, 1:]
                assert np.array_equal(v, want), ("view", tid, it)
                parts = ta.split(0)
                acc = parts[0]
                for p in parts[1:]:
                    acc = acc + p
                assert np.array_equal(acc.flatArray(), a.astype(np.int64).sum(axis=0).astype(np.float32)), ("axis", tid, it)
        except BaseException as e:  # noqa: BLE001 — collected and re-raised on the main thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    cuda.synchronize()


def test_event_wait_lists_and_completion_callbacks(cuda):
    L = cuda._L()
    n = 1 << 20
    host_in = np.arange(n, dtype=np.float32) % 251
    host_out = np.zeros(n, np.float32)
    # upload -> (event) -> kernel that waits on it -> (event) -> read-back that waits on it -> callback from a driver thread
    buf, ev_up = C.c_uint64(), C.c_uint64()
    cuda.check(L.cc_buffer_from_host(host_in.ctypes.data, n, C.byref(buf), C.byref(ev_up)))
    src = cuda.Tensor.fromBuffer(cuda.Buffer(buf.value).share(), [n])
    k = (src * cuda.Tensor.fill(2.0, [n]) + cuda.Tensor.fill(1.0, [n])).compile()
    assert k.info.n_args == 1
    out = cuda.Buffer.alloc(n)
    ev_k, ev_down = C.c_uint64(), C.c_uint64()
    args = (C.c_uint64 * 1)(buf.value)
    waits = (C.c_uint64 * 1)(ev_up.value)
    cuda.check(L.cc_launch(k.handle, args, 1, out.handle, waits, 1, C.byref(ev_k)))
    waits2 = (C.c_uint64 * 1)(ev_k.value)
    cuda.check(L.cc_buffer_to_host(out.handle, 0, host_out.ctypes.data, n, waits2, 1, C.byref(ev_down)))
    fired = threading.Event()
    seen = {}

    @C.CFUNCTYPE(None, C.c_void_p, C.c_int)
    def on_done(user, status):
        seen["status"] = status
        seen["thread"] = threading.get_ident()
        fired.set()

    cuda.check(L.cc_event_on_complete(ev_down.value, C.cast(on_done, C.c_void_p), None))
    assert fired.wait(30.0), "completion callback never ran"
    assert seen["status"] == 0 and seen["thread"] != threading.get_ident()
    done = C.c_int()
    cuda.check(L.cc_event_query(ev_down.value, C.byref(done)))
    assert done.value == 1
    assert np.array_equal(host_out, host_in * 2 + 1)
    for e in (ev_up, ev_k, ev_down):
        cuda.check(L.cc_event_release(e.value))
    cuda.check(L.cc_buffer_release(buf.value))
    out.release()


def test_builtin_profiler(cuda):
    """cc_profile_*: per-command device time, aggregated per kernel structure (the reference's queues have no profiling, O:431-436)"""
    import numpy as np

    T = cuda.Tensor
    a, b = T.random([1024, 1024], seed=1).doCache(), T.random([1024, 1024], seed=2).doCache()
    e = T.tanh(a * b)
    e.flatArray()  # compiled and warm
    cuda.profile(True)
    try:
        for _ in range(5):
            e.doBuffer().release()
        s = a.sum().flatArray()
        host = e.flatArray()
        rep = cuda.profile_report()
    finally:
        cuda.profile(False)
    by = {r["name"].split(":")[0].split(" #")[0]: r for r in rep}
    ew = by["elementwise"]
    assert ew["count"] == 6 and ew["algorithmic_bytes"] == 3 * 4 * 1024 * 1024
    assert 1.0 < ew["avg_us"] < 200.0 and ew["min_us"] <= ew["avg_us"] <= ew["max_us"] and ew["GBs"] > 50.0
    assert "dims=[1024,1024]" in [r["name"] for r in rep if r["name"].startswith("elementwise")][0]
    # the 1-float sum is stored into host memory by its own kernel; only the 4 MiB read-back is a copy command
    assert by["copy device -> host"]["count"] == 1 and by["copy device -> host"]["algorithmic_bytes"] == 4 * 1024 * 1024
    assert any(k.startswith("sum") or k.startswith("whole-tensor fold") for k in by)
    assert np.isfinite(s).all() and host.shape == (1024 * 1024,)
    assert cuda.profile_report() == []  # consumed; nothing recorded while off
    e.doBuffer().release()
    assert cuda.profile_report() == []


def test_c_consumer_runs_config_1(cuda, tmp_path):
    """examples/c1_from_c.c: BASELINE config 1 built and evaluated from plain C through include/compute_cuda.h, in its own process"""
    import re
    import subprocess

    from test_abi_and_codegen import _build_c_consumer

    exe = _build_c_consumer(tmp_path)
    r = subprocess.run([exe, "3000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    m = re.search(r"C1 device-resident from C: ([0-9.]+) us/step", r.stdout)
    assert m and 0.5 < float(m.group(1)) < 100.0, r.stdout
    assert "compiles=" in r.stdout and "max relative error" in r.stdout
    print(r.stdout)


def test_graph_replay_matches_direct_evaluation_and_follows_its_inputs(cuda):
    """cc_graph_*: evaluations recorded once and replayed as one CUDA graph give the same bits as direct evaluation, and a replay reads
    the inputs' CURRENT contents (the graph bakes addresses, not values)"""
    T = cuda.Tensor
    n = 512
    rng = np.random.default_rng(4)
    ha, hb, hc = (rng.standard_normal((n, n)).astype(np.float32) for _ in range(3))
    ba = cuda.Buffer.from_host(ha)
    a, b, c = T.fromBuffer(ba, [n, n]), T(hb), T(hc)
    e = T.tanh(a * b + c)
    col = a.split(0)[0]
    for p in a.split(0)[1:]:
        col = col + p
    direct, direct_col = e.flatArray(), col.flatArray()
    with cuda.Graph() as g:
        for _ in range(10):
            e.doBuffer().release()
        out = e.doBuffer()
        out_col = col.doBuffer()
        total = (a * b).sum().doBuffer()
    assert g.commands >= 13
    g.launch()
    assert np.array_equal(out.to_host(n * n).view(np.uint32), direct.view(np.uint32))
    assert np.array_equal(out_col.to_host(n), direct_col)
    t1 = total.to_host(1)[0]
    assert abs(t1 - float((ha.astype(np.float64) * hb).sum())) <= 1e-3 * n
    # new input contents, same buffers: the replay recomputes from them
    ha2 = (ha * 0.5 + 1.0).astype(np.float32)
    ba.upload(ha2.ctypes.data, n * n)
    for _ in range(3):
        g.launch()
    a2 = T(ha2)
    assert np.array_equal(out.to_host(n * n).view(np.uint32), T.tanh(a2 * b + c).flatArray().view(np.uint32))
    g.release()
    for x in (out, out_col, total, ba):
        x.release()
    # the runtime launches normally again
    assert np.array_equal(e.flatArray().view(np.uint32), T.tanh(a2 * b + c).flatArray().view(np.uint32))
