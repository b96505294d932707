"""Pins the CPU oracle (oracle/reference.py) against every golden vector the reference's own tests
hold for the hot path.  Citations are file:line under /root/reference."""
import numpy as np
import pytest

from oracle import reference as ref
from oracle.reference import Tensor

TS = "Tensors/src/test/scala/com/thoughtworks/compute/TensorsSpec.scala"


def test_tensor_literal():  # TensorsSpec.scala:57-65
    assert str(Tensor(42.0)) == "42.0"
    assert str(Tensor([1.0, 2.0])) == "[1.0,2.0]"
    assert str(Tensor([[1.0, 2.0], [3.0, 4.0]])) == "[[1.0,2.0],[3.0,4.0]]"


def test_wrong_tensor_shape():  # TensorsSpec.scala:67-73
    with pytest.raises(ValueError):
        Tensor([[1.0], [3.0, 4.0]])


def test_fill():  # TensorsSpec.scala:37-55
    t = Tensor.fill(42.0, [2, 3, 5])
    a = t.flat_array()
    assert a.size == 30 and (a == 42.0).all()


def test_translate_with_padding():  # TensorsSpec.scala:75-113
    t = Tensor.fill(42.0, [2, 3, 5], padding=99.0).translate([1, 2, -3])
    assert str(t) == (
        "[[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0]],"
        "[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[42.0,42.0,99.0,99.0,99.0]]]"
    )
    a = t.flat_array().reshape(2, 3, 5)
    for i in range(2):
        for j in range(3):
            for k in range(5):
                assert a[i, j, k] == (42.0 if (i >= 1 and j >= 2 and 5 - k > 3) else 99.0)


def test_unzip():  # TensorsSpec.scala:115-121
    t = Tensor([[[[1.0, 5.0]]]])
    assert [str(s) for s in t.split(3)] == ["[[[1.0]]]", "[[[5.0]]]"]


def test_plus_and_times():  # TensorsSpec.scala:123-138
    t = Tensor([[[1.0, 5.0]]])
    assert str(t + t) == "[[[2.0,10.0]]]"
    t2 = t + t
    assert str(t2 * t2) == "[[[4.0,100.0]]]"


def convolute(input, weight, bias):  # TensorsSpec.scala:144-210
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
    return Tensor.join(outs)


def test_convolution():  # TensorsSpec.scala:140-249
    inp = np.zeros((2, 4, 5, 3), np.float32)
    inp[0, 0, 0, 0] = 1.0
    inp[0, 1, 0, 0] = 10.0
    inp[1, 0, 0, 0] = 100.0
    w = np.zeros((3, 3, 3, 2), np.float32)
    w[1, 1, 0, 0] = 3.0
    w[1, 1, 0, 1] = 4.0
    w[0, 1, 0, 0] = 5.0
    w[2, 2, 0, 1] = 6.0
    out = convolute(Tensor(inp), Tensor(w), Tensor([100000.0, 200000.0]))
    assert out.shape == (2, 4, 5, 2)
    o = out.flat_array().reshape(2, 4, 5, 2)
    assert o[0, 0, 0, 0] == 100053.0
    assert o[0, 1, 1, 1] == 200006.0
    assert o[1, 1, 1, 1] == 200600.0
    assert o[0, 2, 1, 1] == 200060.0
    assert o[0, 0, 0, 1] == 200004.0
    assert o[0, 1, 0, 0] == 100030.0
    assert o[1, 0, 0, 0] == 100300.0


def test_sum():  # TensorsSpec.scala:251-257
    assert str(Tensor.fill(15625.0, [8, 8]).sum()) == "1000000.0"


def test_random_golden_bit_exact():  # TensorsSpec.scala:402-409
    assert str(Tensor.random([3, 3], seed=12345)) == (
        "[[0.48931676,0.2949697,0.14271837],[0.9694414,0.26660874,0.07228618],[0.8779875,0.7046564,0.018829918]]"
    )


RANDOM_NORMAL_GOLDEN = [  # TensorsSpec.scala:414-431
    1.4561316, -0.8711971, -0.7223376, -2.232667, -0.24489015, -0.41490105, -1.0286478, -1.392045, 0.08673929,
    -0.37037173, 0.5294154, -0.5261399, -0.88834476, -0.66154, 0.7035836, -1.1797824, -0.93145895, -1.0812063,
    -1.881317, 0.20438789, -2.5961785, 1.3082669, 0.58748704, -0.01997061, -1.7090794, 1.0162057, 0.33355764,
]


def test_random_normal_golden_within_2ulp():  # TensorsSpec.scala:259-265, 411-434
    got = Tensor.randomNormal([3, 3, 3], seed=54321).flat_array()
    want = np.array(RANDOM_NORMAL_GOLDEN, np.float32)
    # sqrt/log/cos/sin come from the OpenCL driver's libm in the reference: <= 2 ulp of slack per value
    # (cos/sin results near 0 amplify the argument's rounding; bound the absolute error too).
    d = ref.ulp_distance(got, want)
    assert ((d <= 2) | (np.abs(got - want) <= 2e-7)).all(), (got, want, d)
    s = Tensor.randomNormal([], seed=54321).flat_array()
    assert s.shape == (1,) and ref.ulp_distance(s, np.array([1.4561316], np.float32))[0] <= 2


def test_transpose():  # TensorsSpec.scala:436-466
    assert str(Tensor(42.0).transpose()) == "42.0"
    assert str(Tensor([1.0, 2.0, 3.0]).transpose()) == "[1.0,2.0,3.0]"
    assert str(Tensor([[1.0, 2.0], [3.0, 4.0]]).transpose()) == "[[1.0,3.0],[2.0,4.0]]"
    t = Tensor([[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [[7.0, 8.0, 9.0], [10.0, 11.0, 12.0]]])
    assert str(t.transpose()) == "[[[1.0,7.0],[4.0,10.0]],[[2.0,8.0],[5.0,11.0]],[[3.0,9.0],[6.0,12.0]]]"


M1 = [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]
M2 = [[7.0, 8.0, 9.0, 10.0], [11.0, 12.0, 13.0, 14.0], [15.0, 16.0, 17.0, 18.0]]
MM = "[[74.0,80.0,86.0,92.0],[173.0,188.0,203.0,218.0]]"


def matmul2(m1, m2):  # TensorsSpec.scala:472-479 ; benchmarks.scala:188-191
    i, j = m1.shape
    j2, k = m2.shape
    assert j == j2
    product = m1.broadcast([i, j, k]) * m2.reshape([1, j, k]).broadcast([i, j, k])
    terms = product.split(1)
    acc = terms[0]
    for t in terms[1:]:
        acc = acc + t
    return acc


def matmul1(m1, m2):  # TensorsSpec.scala:506-518 ; benchmarks.scala:178-186
    cols1 = m1.split(1)
    out = []
    for col2 in m2.split(1):
        terms = [l * r.broadcast(l.shape) for l, r in zip(cols1, col2.split(0))]
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        out.append(acc)
    return Tensor.join(out)


def test_matrix_multiplication():  # TensorsSpec.scala:468-489
    assert str(matmul2(Tensor(M1), Tensor(M2))) == MM


def test_unrolled_matrix_multiplication():  # TensorsSpec.scala:502-528
    assert str(matmul1(Tensor(M1), Tensor(M2))) == MM


def test_broadcast_trailing():  # TensorsSpec.scala:491-500
    assert str(Tensor(M1).broadcast([2, 3, 4])) == (
        "[[[1.0,1.0,1.0,1.0],[2.0,2.0,2.0,2.0],[3.0,3.0,3.0,3.0]],[[4.0,4.0,4.0,4.0],[5.0,5.0,5.0,5.0],[6.0,6.0,6.0,6.0]]]"
    )


def test_cpu_scaladoc_examples():  # cpu/src/main/scala/com/thoughtworks/compute/cpu.scala:15-101
    assert str(Tensor([[1.0, 2.0], [3.0, 4.0]])) == "[[1.0,2.0],[3.0,4.0]]"
    t = Tensor(np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    s0 = t.split(0)
    assert [str(x) for x in s0] == [
        "[[0.0,1.0,2.0,3.0],[4.0,5.0,6.0,7.0],[8.0,9.0,10.0,11.0]]",
        "[[12.0,13.0,14.0,15.0],[16.0,17.0,18.0,19.0],[20.0,21.0,22.0,23.0]]",
    ]
    assert all(x.shape == (3, 4) for x in s0)
    s1 = t.split(1)
    assert [str(x) for x in s1] == [
        "[[0.0,1.0,2.0,3.0],[12.0,13.0,14.0,15.0]]",
        "[[4.0,5.0,6.0,7.0],[16.0,17.0,18.0,19.0]]",
        "[[8.0,9.0,10.0,11.0],[20.0,21.0,22.0,23.0]]",
    ]
    merged = Tensor.join(
        [Tensor([[1.0, 2.0], [3.0, 4.0]]), Tensor([[5.0, 6.0], [7.0, 8.0]]), Tensor([[9.0, 10.0], [11.0, 12.0]])]
    )
    assert str(merged) == "[[[1.0,5.0,9.0],[2.0,6.0,10.0]],[[3.0,7.0,11.0],[4.0,8.0,12.0]]]"
    assert merged.shape == (2, 2, 3)
    assert str(Tensor.scalar(42.0).broadcast([2, 3])) == "[[42.0,42.0,42.0],[42.0,42.0,42.0]]"


def test_cpu_spec():  # cpu/src/test/scala/com/thoughtworks/compute/cpuSpec.scala:9-38
    a = Tensor.fill(2.0, [2, 3]).non_inline()
    b = Tensor.fill(2.0, [2, 3]).non_inline()
    c = (a + b).non_inline()
    d = (c + b).non_inline()
    assert str(d) == "[[6.0,6.0,6.0],[6.0,6.0,6.0]]"
    a, b = Tensor.fill(42.0, [3, 4]), Tensor.fill(43.0, [3, 4])
    r42, r43 = "[42.0,42.0,42.0,42.0]", "[43.0,43.0,43.0,43.0]"
    t0 = Tensor.join([a, b], 0)
    assert t0.shape == (2, 3, 4)
    assert str(t0) == "[[%s],[%s]]" % (",".join([r42] * 3), ",".join([r43] * 3))
    t1 = Tensor.join([a, b], 1)
    assert t1.shape == (3, 2, 4)
    assert str(t1) == "[" + ",".join(["[%s,%s]" % (r42, r43)] * 3) + "]"
    t2 = Tensor.join([a, b], 2)
    assert t2.shape == (3, 4, 2)
    assert str(t2) == "[" + ",".join(["[" + ",".join(["[42.0,43.0]"] * 4) + "]"] * 3) + "]"


def test_affine_concatenate_matches_awt_law():
    """NDimensionalAffineTransformSpec.scala:18-56: preConcatenate/concatenate agree with
    java.awt.geom.AffineTransform, i.e. with 3x3 homogeneous matrix products."""

    def h(m):
        return np.array([m[0:3], m[3:6], [0, 0, 1]], dtype=np.float64)

    rng = np.random.RandomState(0)
    cases = [([1.0, 0.0, 3.5, 0.0, 1.0, 4.2], [3.0, 0.0, 0.0, 0.0, 2.0, 0.0])]
    cases.append((list(rng.randint(0, 100, 6).astype(float)), list(rng.randint(0, 100, 6).astype(float))))
    for m0, m1 in cases:
        # at = m0; at.preConcatenate(m1)  =>  m1 * m0
        got = ref.nd_pre_concatenate(m0, m1, 2)
        assert np.array_equal(h(got), h(m1) @ h(m0))
        # at = m1; at.concatenate(m0)  =>  m1 * m0
        got = ref.nd_concatenate(m1, m0, 2)
        assert np.array_equal(h(got), h(m1) @ h(m0))


def test_decimal_format_and_truncation_quirks():
    """OpenCLKernelBuilder.scala:14-32,372-379,386 — 3 fraction digits, (int) truncates toward zero."""
    assert ref.java_decimal_format(1.0 / 3) == "0.333"
    assert ref.java_decimal_format(5.0) == "5"
    assert ref.java_decimal_format(-7.0) == "-7"
    assert ref.java_decimal_format(2.5) == "2.5"
    assert ref.java_decimal_format(0.0625) == "0.062"  # HALF_EVEN
    src = np.arange(3, dtype=np.float32)
    # scale a length-3 vector to length 9: index = (int)(gid * 0.333)
    got = ref.affine_gather(src, [1.0 / 3, 0.0], (9,), 0.0)
    assert got.tolist() == [0, 0, 0, 0, 1, 1, 1, 2, 2]
    # an index in (-1, 0) truncates to 0 and is therefore in range
    got = ref.affine_gather(src, [0.5, -0.5], (3,), 99.0)
    assert got.tolist() == [0.0, 0.0, 0.0]


def test_sum_reference_order_matches_definition():
    """Tensors.scala:313-351 with global size 1 — 16 sequential lanes, hi/lo tree, tail."""
    rng = np.random.RandomState(1)
    x = rng.rand(16 * 7 + 5).astype(np.float32)
    lanes = x[:16].copy()
    for v in range(1, 7):
        lanes = (lanes + x[16 * v : 16 * v + 16]).astype(np.float32)
    f8 = lanes[8:] + lanes[:8]
    f4 = f8[4:] + f8[:4]
    f2 = f4[2:] + f4[:2]
    s = np.float32(f2[0] + f2[1])
    for v in x[112:]:
        s = np.float32(s + v)
    assert ref.sum_reference_cpu_order(x) == s
    y = rng.rand(9).astype(np.float32)
    s = np.float32(0)
    for v in y:
        s = np.float32(s + v)
    assert ref.sum_reference_cpu_order(y) == s
