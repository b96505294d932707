"""Pins the C/OpenMP port (oracle/oracle_cpu.c — the CPU baseline bench.py times) to the numpy oracle, which is pinned to the
reference's golden vectors.  CPU only."""
import numpy as np
import pytest

from oracle import build as ob
from oracle import reference as ref


@pytest.fixture(scope="module", params=["strict", "fma"])
def L(request):
    ob.build()
    return ob.load(request.param), request.param


def test_random_matches_bit_exact(L):
    lib, _ = L
    for n, seed in ((9, 12345), (1000, 7), (4097, 0xFFFFFFFF)):
        out = np.empty(n, np.float32)
        lib.oracle_random(out.ctypes.data, n, seed)
        assert np.array_equal(out.view(np.uint32), ref.random_buffer(n, seed).view(np.uint32))
    out = np.empty(9, np.float32)
    lib.oracle_random(out.ctypes.data, 9, 12345)  # TensorsSpec.scala:405-406
    assert [ref.java_float_to_string(v) for v in out[:3]] == ["0.48931676", "0.2949697", "0.14271837"]


def test_c1_c2_within_libm_slack(L):
    lib, variant = L
    n = 1 << 16
    a, b, c = (ref.random_buffer(n, s) for s in (1, 2, 3))
    out = np.empty(n, np.float32)
    lib.oracle_c1(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n)
    R = ref.Tensor
    ref.CONTRACT[0] = variant == "fma"
    try:
        want = R.tanh(R(a) * R(b) + R(c)).flat_array()
    finally:
        ref.CONTRACT[0] = False
    assert ref.ulp_distance(out, want).max() <= 2  # glibc tanhf (<= 2 ulp) vs correctly rounded
    lib.oracle_c2(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n)
    a64, b64, c64 = (x.astype(np.float64) for x in (a, b, c))
    truth = np.tanh(np.log(np.exp(a64 * b64 + c64) + a64) * b64) + c64
    assert np.abs(out - truth).max() <= 4e-6  # ill-conditioned chain: absolute agreement only (tests/test_parity_configs.py has the bound)


def test_sum_order_and_axis_sums(L):
    lib, _ = L
    for n in (5, 16, 100, 4099):
        x = ref.random_buffer(n, 5)
        assert np.float32(lib.oracle_sum_cpu_order(x.ctypes.data, n)) == ref.sum_reference_cpu_order(x)  # Tensors.scala:313-351
    rows, cols = 37, 53
    x = ref.random_buffer(rows * cols, 5).reshape(rows, cols)
    for axis in (0, 1):
        out = np.empty(cols if axis == 0 else rows, np.float32)
        lib.oracle_axis_sum_2d(x.ctypes.data, rows, cols, axis, out.ctypes.data)
        want = np.add.accumulate(np.moveaxis(x, axis, 0), axis=0, dtype=np.float32)[-1]
        assert np.array_equal(out, want)


def test_matmul_left_fold_and_gather(L):
    lib, variant = L
    m, k, n = 9, 13, 7
    a = (np.floor(ref.random_buffer(m * k, 9) * 9) - 4).astype(np.float32).reshape(m, k)
    b = (np.floor(ref.random_buffer(k * n, 10) * 9) - 4).astype(np.float32).reshape(k, n)
    c = np.empty((m, n), np.float32)
    lib.oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n)
    assert np.array_equal(c, (a.astype(np.int64) @ b.astype(np.int64)).astype(np.float32))
    # TensorsSpec.scala:485 golden
    a = np.asarray([[1, 2, 3], [4, 5, 6]], np.float32)
    b = np.asarray([[7, 8, 9, 10], [11, 12, 13, 14], [15, 16, 17, 18]], np.float32)
    c = np.empty((2, 4), np.float32)
    lib.oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, c.ctypes.data, 2, 3, 4)
    assert c.tolist() == [[74, 80, 86, 92], [173, 188, 203, 218]]
    # affine gather vs the numpy oracle: permute(2,0,1) + translate(3,-5,7)
    d = 12
    t = ref.random_buffer(d**3, 7).reshape(d, d, d)
    want = ref.Tensor(t).permute([2, 0, 1]).translate([3, -5, 7]).flat_array()
    mat = np.asarray([[0, 1, 0, 5], [0, 0, 1, -7], [1, 0, 0, -3]], np.int64)
    out = np.empty(d**3, np.float32)
    shp = np.asarray([d, d, d], np.int64)
    lib.oracle_affine_gather_3d(t.ctypes.data, shp.ctypes.data, 3, mat.ctypes.data, shp.ctypes.data, 0.0, out.ctypes.data)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the generated kernel on the host cores): one JSON line with the keys the driver
    reads; run here on a small sample so that the contract is checked without a GPU"""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "3"], capture_output=True, text=True, cwd=root,
                       env=dict(os.environ, BENCH_REFERENCE_SAMPLE_LOG2="20", OMP_NUM_THREADS="1"), timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    # torchrun exports OMP_NUM_THREADS=1 to every rank (as this test does): the reference arm still uses every host thread
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["workload"].startswith("C2 long fused elementwise chain") and d["config"]["elements_per_gpu"] == 1 << 20
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["metric"].startswith("fused elementwise HBM GB/s")
