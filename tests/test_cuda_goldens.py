"""GPU parity, part 1: the reference's own test-suite re-run against `cuda` through the C ABI.

Every case mirrors a case of the reference (file:line cited, paths under /root/reference) and asserts the same
golden value the reference asserts; in addition each result is compared with the CPU oracle on the same inputs.
"""
import numpy as np
import pytest

from oracle import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    from compute.scala_b200 import cuda as c

    c.init()
    yield c
    c.synchronize()


def both(cuda):
    return cuda.Tensor, ref.Tensor


def test_tensor_literal(cuda):  # TensorsSpec.scala:57-65
    T = cuda.Tensor
    assert str(T(42.0)) == "42.0"
    assert str(T([1.0, 2.0])) == "[1.0,2.0]"
    assert str(T([[1.0, 2.0], [3.0, 4.0]])) == "[[1.0,2.0],[3.0,4.0]]"


def test_wrong_tensor_shape(cuda):  # TensorsSpec.scala:67-73
    with pytest.raises(ValueError):
        cuda.Tensor([[1.0], [3.0, 4.0]])


def test_fill_and_kernel_cache(cuda):  # TensorsSpec.scala:37-55
    T = cuda.Tensor
    t = T.fill(42.0, [2, 3, 5])
    a = t.flatArray()
    assert a.size == 30 and (a == 42.0).all()
    before = cuda.stats()
    t2 = T.fill(42.0, [2, 3, 5])  # structurally equal closure -> served from the cache (:50-52)
    assert (t2.flatArray() == 42.0).all()
    after = cuda.stats()
    assert after["compiles"] == before["compiles"]
    assert after["cache_hits"] == before["cache_hits"] + 1


def test_translate_with_padding(cuda):  # TensorsSpec.scala:75-113
    t = cuda.Tensor.fill(42.0, [2, 3, 5], padding=99.0).translate([1, 2, -3])
    assert str(t) == (
        "[[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0]],"
        "[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[42.0,42.0,99.0,99.0,99.0]]]"
    )


def test_unzip(cuda):  # TensorsSpec.scala:115-121
    t = cuda.Tensor([[[[1.0, 5.0]]]])
    assert [str(s) for s in t.split(3)] == ["[[[1.0]]]", "[[[5.0]]]"]


def test_plus_and_times(cuda):  # TensorsSpec.scala:123-138
    t = cuda.Tensor([[[1.0, 5.0]]])
    assert str(t + t) == "[[[2.0,10.0]]]"
    t2 = t + t
    assert str(t2 * t2) == "[[[4.0,100.0]]]"


def convolute(T, input, weight, bias):  # TensorsSpec.scala:144-210
    batch, height, width, depth = input.shape
    kh, kw, depth2, filters = weight.shape
    assert depth2 == depth and bias.shape == (filters,)
    input_seq = input.split(3)
    weight_seq = [[[d.split(0) for d in kwd.split(0)] for kwd in khkwd.split(0)] for khkwd in weight.split(3)]
    bias_seq = bias.split(0)
    outs = []
    for w_f, b_f in zip(weight_seq, bias_seq):
        summands = []
        for oy, w_row in zip((-1, 0, 1), w_f):
            for ox, w_px in zip((-1, 0, 1), w_row):
                for in_c, w_c in zip(input_seq, w_px):
                    assert w_c.shape == ()
                    summands.append(in_c.translate([0, oy, ox]) * w_c.broadcast([batch, height, width]))
        acc = summands[0]
        for s in summands[1:]:
            acc = acc + s
        outs.append(b_f.broadcast([batch, height, width]) + acc)
    return T.join(outs)


def test_convolution(cuda):  # TensorsSpec.scala:140-249 — the reference's spot values, then a dense case vs the oracle
    inp = np.zeros((2, 4, 5, 3), np.float32)
    inp[0, 0, 0, 0] = 1.0
    inp[0, 1, 0, 0] = 10.0
    inp[1, 0, 0, 0] = 100.0
    w = np.zeros((3, 3, 3, 2), np.float32)
    w[1, 1, 0, 0] = 3.0
    w[1, 1, 0, 1] = 4.0
    w[0, 1, 0, 0] = 5.0
    w[2, 2, 0, 1] = 6.0
    T = cuda.Tensor
    out = convolute(T, T(inp), T(w), T([100000.0, 200000.0]))
    assert out.shape == (2, 4, 5, 2)
    o = out.flatArray().reshape(2, 4, 5, 2)
    assert o[0, 0, 0, 0] == 100053.0
    assert o[0, 1, 1, 1] == 200006.0
    assert o[1, 1, 1, 1] == 200600.0
    assert o[0, 2, 1, 1] == 200060.0
    assert o[0, 0, 0, 1] == 200004.0
    assert o[0, 1, 0, 0] == 100030.0
    assert o[1, 0, 0, 0] == 100300.0
    rng = np.random.default_rng(0)
    inp = rng.integers(-3, 4, (2, 6, 7, 3)).astype(np.float32)
    w = rng.integers(-3, 4, (3, 3, 3, 2)).astype(np.float32)
    b = np.asarray([0.5, -1.5], np.float32)
    got = convolute(T, T(inp), T(w), T(b)).flatArray()
    R = ref.Tensor
    want = convolute(R, R(inp), R(w), R(b)).flat_array()
    assert got.view(np.uint32).tolist() == want.view(np.uint32).tolist()  # small integers: exact in any order


def test_sum(cuda):  # TensorsSpec.scala:251-257
    t = cuda.Tensor.fill(15625.0, [8, 8])
    assert str(t.sum()) == "1000000.0"


RANDOM_NORMAL_GOLDEN = [  # TensorsSpec.scala:414-431
    1.4561316, -0.8711971, -0.7223376, -2.232667, -0.24489015, -0.41490105, -1.0286478, -1.392045, 0.08673929,
    -0.37037173, 0.5294154, -0.5261399, -0.88834476, -0.66154, 0.7035836, -1.1797824, -0.93145895, -1.0812063,
    -1.881317, 0.20438789, -2.5961785, 1.3082669, 0.58748704, -0.01997061, -1.7090794, 1.0162057, 0.33355764,
]


def test_random(cuda):  # TensorsSpec.scala:402-409 — bit-exact
    assert str(cuda.Tensor.random([3, 3], seed=12345)) == (
        "[[0.48931676,0.2949697,0.14271837],[0.9694414,0.26660874,0.07228618],[0.8779875,0.7046564,0.018829918]]"
    )
    for n, seed in ((1, 0), (7, -5), (4096 + 3, 99), (1 << 20, 2**31 - 1)):
        got = cuda.Tensor.random([n], seed=seed).flatArray()
        want = ref.random_buffer(n, seed)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_random_normal(cuda):  # TensorsSpec.scala:259-265, 411-434 — sqrt/log/cos/sin come from the device libm
    got = cuda.Tensor.randomNormal([3, 3, 3], seed=54321).flatArray()
    want = np.array(RANDOM_NORMAL_GOLDEN, np.float32)
    d = ref.ulp_distance(got, want)
    assert ((d <= 2) | (np.abs(got - want) <= 2e-7)).all(), (got, want, d)
    odd = cuda.Tensor.randomNormal([5], seed=7).flatArray()  # odd size: last pair half-written (:509-513)
    want_odd = ref.random_normal_buffer(5, 7)
    d = ref.ulp_distance(odd, want_odd)
    assert ((d <= 4) | (np.abs(odd - want_odd) <= 4e-7)).all()


def test_random_normal_large_and_its_singular_pair(cuda):  # Tensors.scala:398-429 at benchmark sizes (benchmarks.scala:372-376)
    """hash(61) == 0 (Tensors.scala:106-117), so the pair with index 61 ^ seed draws u1 = 0: r = sqrt(-2 log 0) = +inf,
    theta = 0 -> z0 = +inf, z1 = inf * 0 = NaN.  The reference produces exactly that; so must we, and nothing else non-finite."""
    n, seed = 1 << 16, 7
    got = cuda.Tensor.randomNormal([n], seed=seed).flatArray()
    want = ref.random_normal_buffer(n, seed)
    p = 61 ^ seed
    assert np.isposinf(got[2 * p]) and np.isnan(got[2 * p + 1]) and np.isposinf(want[2 * p]) and np.isnan(want[2 * p + 1])
    fin = np.ones(n, bool)
    fin[2 * p : 2 * p + 2] = False
    assert np.isfinite(got[fin]).all() and np.isfinite(want[fin]).all()
    d = ref.ulp_distance(got[fin], want[fin])
    # cos / sin of a large-ish theta near a zero crossing: compare absolutely there (|z| <= r * 4 ulp(theta))
    assert ((d <= 4) | (np.abs(got[fin] - want[fin]) <= 4e-6)).all(), d.max()
    assert abs(float(got[fin].mean())) < 0.02 and abs(float(got[fin].std()) - 1.0) < 0.02


def test_flat_buffer_is_pinned_and_pooled(cuda):  # flatBuffer, Tensors.scala:1099-1109; host memory O:691-715
    T = cuda.Tensor
    e = T.tanh(T.random([257, 129], seed=3) * T.fill(2.0, [257, 129]))
    want = e.flatArray()
    hb = e.flatBuffer()
    assert hb.n == 257 * 129 and np.array_equal(hb.array.view(np.uint32), want.view(np.uint32))
    first = hb.array.ctypes.data
    hb.release()
    hb.release()  # idempotent
    with e.flatBuffer() as a:  # the block comes back from the pool: no new cuMemHostAlloc
        assert a.ctypes.data == first and np.array_equal(a.view(np.uint32), want.view(np.uint32))
    with T.scalar(3.0).flatBuffer() as a:
        assert a.tolist() == [3.0]
    # freeing memory the library did not hand out is an error, not a crash
    import ctypes as C

    assert cuda._L().cc_host_free(C.c_void_p(0x1000)) == -1


def test_small_results_are_stored_into_host_memory_by_the_kernel(cuda):
    """flatArray / flatBuffer of <= 16384 floats: the kernel's output buffer IS pinned host memory (no copy command); the values
    are those of the ordinary device-buffer route"""
    T = cuda.Tensor
    a, b = T.random([96, 128], seed=1).doCache(), T.random([96, 128], seed=2).doCache()
    cases = {
        "elementwise": T.tanh(a * b) + a,
        "view": (a * b).permute([1, 0]).translate([1, -1]),
        "transpose": a.transpose(),
        "fold": (a * b).sum(),
        "sum of a buffer": a.sum(),
        "max": a.reduce("max"),
        "join": T.join([a, b, a * b]).split(0)[3],
        "axis sum": _chain(a.split(0)),
        "small matmul (generic reduction)": _chain((a.broadcast([96, 128, 128]) * T.random([128, 128], seed=3).reshape([1, 128, 128]).broadcast([96, 128, 128])).split(1)),
        "matmul (tensor cores: not redirected)": _chain(
            (T.random([128, 2048], seed=5).broadcast([128, 2048, 128]) * T.random([2048, 128], seed=6).reshape([1, 2048, 128]).broadcast([128, 2048, 128])).split(1)
        ),
        "random": T.random([100], seed=9),
        "scalar": T.scalar(2.5) * T.scalar(4.0),
    }
    for name, e in cases.items():
        buf = e.doBuffer()
        n = int(np.prod(e.shape)) if e.shape else 1
        via_device = buf.to_host(n)
        buf.release()
        before = cuda.stats()["d2h_bytes"]
        got = e.flatArray()
        with e.flatBuffer() as pinned:
            assert np.array_equal(pinned.view(np.uint32), via_device.view(np.uint32)), name
        copied = cuda.stats()["d2h_bytes"] - before
        assert np.array_equal(got.view(np.uint32), via_device.view(np.uint32)), name
        if name in ("elementwise", "view", "transpose", "fold", "sum of a buffer", "max", "axis sum", "scalar", "small matmul (generic reduction)"):
            assert copied == 0, (name, copied)  # no copy command was issued
        elif name in ("random",) or name.startswith("matmul"):
            assert copied == 2 * 4 * n, (name, copied)
    big = T.tanh(T.random([129, 128], seed=4))  # 16512 floats: over the threshold, ordinary route
    before = cuda.stats()["d2h_bytes"]
    big.flatArray()
    assert cuda.stats()["d2h_bytes"] - before == 4 * 129 * 128


def _chain(parts):
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def test_transpose(cuda):  # TensorsSpec.scala:436-466
    T = cuda.Tensor
    assert str(T.scalar(42.0).transpose()) == "42.0"
    assert str(T([1.0, 2.0, 3.0]).transpose()) == "[1.0,2.0,3.0]"
    assert str(T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]).transpose()) == "[[1.0,4.0],[2.0,5.0],[3.0,6.0]]"
    t3 = T(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    want = ref.Tensor(np.arange(24, dtype=np.float32).reshape(2, 3, 4)).transpose()
    assert str(t3.transpose()) == str(want)
    assert t3.transpose().shape == (4, 3, 2)


def matrix_multiply2(T, m1, m2):  # TensorsSpec.scala:472-479, benchmarks.scala:188-191
    i, j = m1.shape
    j2, k = m2.shape
    assert j == j2
    product = m1.broadcast([i, j, k]) * m2.reshape([1, j, k]).broadcast([i, j, k])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def matrix_multiply1(T, m1, m2):  # TensorsSpec.scala:506-518, benchmarks.scala:178-186
    i, j = m1.shape
    _, k = m2.shape
    cols1 = m1.split(1)
    outs = []
    for column2 in m2.split(1):
        terms = [c1 * s.broadcast([i]) for c1, s in zip(cols1, column2.split(0))]
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        outs.append(acc)
    return T.join(outs)


def test_matrix_multiply(cuda):  # TensorsSpec.scala:468-489, 502-528
    T = cuda.Tensor
    m1 = T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    m2 = T([[7.0, 8.0, 9.0, 10.0], [11.0, 12.0, 13.0, 14.0], [15.0, 16.0, 17.0, 18.0]])
    want = "[[74.0,80.0,86.0,92.0],[173.0,188.0,203.0,218.0]]"
    assert str(matrix_multiply2(T, m1, m2)) == want
    assert str(matrix_multiply1(T, m1, m2)) == want


def test_broadcast(cuda):  # TensorsSpec.scala:491-500
    t = cuda.Tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]).broadcast([2, 3, 4])
    assert str(t) == (
        "[[[1.0,1.0,1.0,1.0],[2.0,2.0,2.0,2.0],[3.0,3.0,3.0,3.0]],[[4.0,4.0,4.0,4.0],[5.0,5.0,5.0,5.0],[6.0,6.0,6.0,6.0]]]"
    )
    with pytest.raises(ValueError):
        cuda.Tensor([[1.0, 2.0, 3.0]]).broadcast([2, 4])


def test_cpu_spec_chain_and_join(cuda):  # cpuSpec.scala:9-38
    T = cuda.Tensor
    a = T.fill(1.0, [2, 2]).nonInline()
    b = (a + a).nonInline()
    c = (b + a + a + a + a).nonInline()
    assert (c.flatArray() == 6.0).all()
    ts = [T(np.full((2, 3), float(v), dtype=np.float32)) for v in (1, 2, 3, 4)]
    rs = [ref.Tensor(np.full((2, 3), float(v), dtype=np.float32)) for v in (1, 2, 3, 4)]
    for d in (0, 1, 2):
        assert str(T.join(ts, d)) == str(ref.Tensor.join(rs, d))


def test_scaladoc_examples(cuda):  # cpu.scala:15-101
    T = cuda.Tensor
    iota = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    t, r = T(iota), ref.Tensor(iota)
    for d in (0, 1, 2):
        assert [str(x) for x in t.split(d)] == [str(x) for x in r.split(d)]
    assert str(T.join(t.split(1))) == str(ref.Tensor.join(r.split(1)))
    assert str(T.scalar(3.0).broadcast([2, 2]) + T([[1.0, 2.0], [3.0, 4.0]])) == "[[4.0,5.0],[6.0,7.0]]"


def test_auto_broadcast_errors(cuda):  # Tensors.scala:208-222
    T = cuda.Tensor
    with pytest.raises(ValueError):
        (T.fill(1.0, [2, 3]) + T.fill(1.0, [3, 3])).flatArray()
    assert (T.fill(1.0, [2, 1]) + T.fill(2.0, [1, 3])).shape == (2, 3)
    assert (T.fill(1.0, [2]) + T.fill(2.0, [2, 3])).shape == (2, 3)  # leading-aligned (finding 7)


def test_no_leaked_buffers(cuda):
    T = cuda.Tensor
    cuda.synchronize()
    base = cuda.stats()["bytes_in_use"]
    for _ in range(3):
        x = T.random([64, 64], seed=1)
        y = T.tanh(x * x + x).flatArray()
        assert y.shape == (4096,)
    del x
    assert cuda.stats()["bytes_in_use"] == base
