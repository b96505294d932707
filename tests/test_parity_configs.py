"""GPU parity, part 2: the five BASELINE.json configurations, CUDA path (through the C ABI) vs the CPU oracle on the
same seeded inputs, at sizes the oracle finishes in seconds; full-size runs are checked through size-independent
properties (tests/test_full_size.py).  Tolerances are the north star's: bit-exact for index/view work, <= 2 ulp for
fused elementwise fp32, <= 1e-5 relative for reductions and matmul."""
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


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def oracle_bracket(build, *args):
    """oracle result with and without FP_CONTRACT (the reference compiles with -cl-unsafe-math-optimizations)"""
    out = []
    for contract in (False, True):
        ref.CONTRACT[0] = contract
        try:
            out.append(build(ref.Tensor, *args).flat_array())
        finally:
            ref.CONTRACT[0] = False
    return out


def min_ulp(got, candidates):
    d = None
    for c in candidates:
        x = ref.ulp_distance(got, c)
        d = x if d is None else np.minimum(d, x)
    return d


# ---- C1: tanh(a*b+c), 1024 x 1024 ------------------------------------------------------------------------------------


def c1(T, n):
    a, b, c = T.random([n, n], seed=1), T.random([n, n], seed=2), T.random([n, n], seed=3)
    return T.tanh(a * b + c)


@pytest.mark.parametrize("n", [1024, 33])
def test_c1_fused_elementwise(cuda, n):
    got = c1(cuda.Tensor, n).flatArray()
    d = min_ulp(got, oracle_bracket(c1, n))
    assert d.max() <= 2, f"max ulp distance {d.max()}"


# ---- C2: long chain with every op kind -----------------------------------------------------------------------------------


def c2(T, shape):
    a, b, c = T.random(shape, seed=1), T.random(shape, seed=2), T.random(shape, seed=3)
    t = a * b + c
    u = T.exp(t)
    v = T.log(u + a)
    w = T.tanh(v * b)
    return w + c


def c2_truth_and_bound(shape, func_ulps=2.0):
    """fp64 evaluation of the chain plus the first-order forward error bound of an fp32 evaluation in which every
    + and * is correctly rounded (relative error <= 2^-24) and exp / log / tanh are within `func_ulps` ulp — i.e. what
    "each fused op within 2 ulp" means for a chain whose intermediate roundings are amplified at ill-conditioned
    points (log(u + a) near u + a = 1)."""
    n = int(np.prod(shape))
    a, b, c = (ref.random_buffer(n, s).astype(np.float64) for s in (1, 2, 3))
    u24, F = 2.0**-24, func_ulps * 2.0**-23
    ab = a * b
    t = ab + c
    e_t = (np.abs(ab) + np.abs(t)) * u24
    u = np.exp(t)
    e_u = u * e_t + u * F
    s_ = u + a
    e_s = e_u + np.abs(s_) * u24
    v = np.log(s_)
    e_v = e_s / s_ + np.abs(v) * F
    p = v * b
    e_p = np.abs(b) * e_v + np.abs(p) * u24
    w = np.tanh(p)
    e_w = (1.0 - w * w) * e_p + np.abs(w) * F
    out = w + c
    e_out = e_w + np.abs(out) * u24
    return out, 1.05 * e_out + 1e-30


@pytest.mark.parametrize("shape", [(1024, 1024), (7, 5, 3), (4099,)])
def test_c2_long_chain(cuda, shape):
    got = c2(cuda.Tensor, list(shape)).flatArray().astype(np.float64)
    truth, bound = c2_truth_and_bound(shape, func_ulps=2.0)
    ratio = np.abs(got - truth) / bound
    assert ratio.max() <= 1.0, f"error is {ratio.max():.2f}x the 2-ulp-per-op bound"
    # the restated reference (fp32 steps, correctly rounded libm) sits inside the same envelope; the two agree to within
    # the sum of their envelopes
    for step in oracle_bracket(c2, list(shape)):
        assert (np.abs(step.astype(np.float64) - truth) <= bound).all()
        assert (np.abs(step.astype(np.float64) - got) <= 2 * bound).all()
    # where the chain is well conditioned (bound below 2 ulp of the result) the plain 2-ulp statement holds
    well = bound <= 2 * np.spacing(np.abs(truth).astype(np.float32)).astype(np.float64)
    if well.any():
        d = ref.ulp_distance(got.astype(np.float32)[well], truth.astype(np.float32)[well])
        assert d.max() <= 2


@pytest.mark.parametrize("op", ["exp", "log", "tanh", "sqrt", "abs"])
def test_single_op_within_2_ulp(cuda, op):
    """each math function alone, fused with the affine map that spreads random() over its interesting domain"""
    T = cuda.Tensor
    n = 1 << 20
    lo, hi = {"exp": (-20.0, 20.0), "log": (1e-6, 30.0), "tanh": (-9.0, 9.0), "sqrt": (0.0, 1e6), "abs": (-1.0, 1.0)}[op]
    x_np = (ref.random_buffer(n, 31) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
    x = T(x_np)
    got = getattr(T, op)(x).flatArray()
    want = getattr(np, op if op != "abs" else "abs")(x_np.astype(np.float64)).astype(np.float32)
    assert ref.ulp_distance(got, want).max() <= 2


def test_all_elementwise_ops(cuda):
    """every node kind of Expressions.scala:103-124 on awkward shapes (ragged sizes exercise the scalar tail path)"""
    for shape in ([5, 7], [64, 64], [3, 1, 9]):

        def build(T):
            x, y = T.random(shape, seed=11), T.random(shape, seed=12)
            one = T.fill(1.0, shape)
            e = T.sqrt(x + one) / (y + one)
            e = T.max(e, x) - T.min(y, e)
            e = T.abs(-e) % (y + one)
            return e * T.exp(x) + T.log(y + one)

        got = build(cuda.Tensor).flatArray()
        d = min_ulp(got, oracle_bracket(build))
        assert d.max() <= 2, (shape, d.max())


def test_iterated_maps_loop_equals_unrolled(cuda, monkeypatch):
    """benchmarks.scala:100-108 / 319-326: the loop the generator emits for a folded map is the unrolled kernel, bit for bit,
    and both sit within the oracle's bracket"""
    T = cuda.Tensor

    def fold(n, x, f):
        for _ in range(n):
            x = f(x)
        return x

    def issue137(T, n, shape):
        a, b, c = (T.random(shape, seed=s) for s in (1, 2, 3))
        return fold(n, a, lambda v: v * b + c)

    def tanh_n(T, n, shape):
        return fold(n, T.random(shape, seed=4) * T.fill(6.0, shape) - T.fill(3.0, shape), T.tanh)

    def mixed(T, n, shape):
        a, b = T.random(shape, seed=1), T.random(shape, seed=2)
        e = fold(n, a, lambda v: T.exp(-v) * v + b)
        return e * e - e

    for build, n, shape in ((issue137, 100, [33, 65]), (tanh_n, 100, [64, 64]), (tanh_n, 10, [7, 5, 3]), (mixed, 30, [128, 36])):
        cuda.kernel_cache_clear()
        k = build(T, n, shape).compile()
        assert "int it_" in k.source
        looped = build(T, n, shape).flatArray()
        monkeypatch.setenv("CC_NO_OP_LOOPS", "1")
        cuda.kernel_cache_clear()
        assert "int it_" not in build(T, n, shape).compile().source
        unrolled = build(T, n, shape).flatArray()
        monkeypatch.delenv("CC_NO_OP_LOOPS")
        cuda.kernel_cache_clear()
        assert np.array_equal(bits(looped), bits(unrolled)), build.__name__
        if n <= 10:
            assert min_ulp(looped, oracle_bracket(build, n, shape)).max() <= 2 * n
    # a contracting affine map converges to its fixed point c / (1 - b) whatever the rounding of each step
    got = issue137(T, 100, [33, 65]).flatArray().astype(np.float64)
    b, c = (ref.random_buffer(33 * 65, s).astype(np.float64) for s in (2, 3))
    fixed = c / (1.0 - b)
    ok = b < 0.8
    assert np.abs(got[ok] - fixed[ok]).max() <= 1e-5 * np.abs(fixed[ok]).max()


# ---- C3: reductions ----------------------------------------------------------------------------------------------------------


def dataset_e(T, shape, seed=5):
    """floor(random*9) - 4 in {-4..4}: exactly summable in any order. floor is built from `%`."""
    r = T.random(shape, seed=seed) * T.fill(9.0, shape)
    return (r - r % T.fill(1.0, shape)) - T.fill(4.0, shape)


def dataset_e_np(n, seed=5):
    r = (ref.random_buffer(n, seed) * np.float32(9.0)).astype(np.float32)
    return (np.floor(r) - np.float32(4.0)).astype(np.float32)


@pytest.mark.parametrize("n", [1024, 1000, 17])
def test_c3_full_sum(cuda, n):
    T = cuda.Tensor
    e = dataset_e(T, [n, n])
    want_e = dataset_e_np(n * n)
    assert np.array_equal(e.flatArray(), want_e)
    got = e.sum().flatArray()[0]
    assert got == ref.sum_reference_cpu_order(want_e) == np.float32(want_e.astype(np.int64).sum())  # bit-exact
    u = T.random([n, n], seed=5)
    got_u = float(u.sum().flatArray()[0])
    truth = float(ref.random_buffer(n * n, 5).astype(np.float64).sum())
    ref_order = float(ref.sum_reference_cpu_order(ref.random_buffer(n * n, 5)))
    assert abs(got_u - truth) <= 1e-5 * abs(truth), (got_u, truth, ref_order)


@pytest.mark.parametrize("shape", [(512, 1024), (1000, 36), (7, 9, 5), (3,), ()])
def test_fused_sum_equals_materialise_then_sum(cuda, shape):
    """`expr.sum` folds the closure of an inline operand inside the reduction kernel (one pass, no intermediate);
    the reference materialises first (Tensors.scala:678). Same fold order => same bits as the two-step sequence."""
    T = cuda.Tensor
    shape = list(shape)
    a, b, c = (T.random(shape, seed=s).doCache() for s in (1, 2, 3))
    expr = T.tanh(a * b + c)
    k = expr.sum().compile()
    assert k.info.kind == 4 and k.info.n_args == 3 and k.info.n_launches == 1
    n = int(np.prod(shape)) if shape else 1
    assert k.info.algorithmic_bytes == 12 * n + 4
    s0 = cuda.stats()
    fused = expr.sum().flatArray()[0]
    s1 = cuda.stats()
    assert s1["device_kernels"] - s0["device_kernels"] == 1  # one kernel: loads, chain and fold
    two_step = expr.doCache().sum().flatArray()[0]
    if n % 4 == 0 and shape and shape[-1] % 4 == 0:
        assert fused.view(np.uint32) == two_step.view(np.uint32)
    want = ref.Tensor.tanh(ref.Tensor.random(shape, seed=1) * ref.Tensor.random(shape, seed=2) + ref.Tensor.random(shape, seed=3)).flat_array()
    truth = float(want.astype(np.float64).sum())
    assert abs(float(fused) - truth) <= 1e-5 * abs(truth)
    assert abs(float(two_step) - truth) <= 1e-5 * abs(truth)
    # exactly summable data: bit-exact in any order
    e = dataset_e(T, shape)
    assert e.sum().flatArray()[0] == np.float32(dataset_e_np(n).astype(np.int64).sum())


def test_other_monoids_and_views_inside_the_fold(cuda):
    """MonoidPrograms is generic over append / zero (Tensors.scala:308-311); only Plus is instantiated by the reference"""
    T = cuda.Tensor
    shape = [37, 50]
    x = ref.random_buffer(37 * 50, 11).reshape(shape)
    t = T.random(shape, seed=11).doCache()
    assert t.reduce("min").flatArray()[0] == x.min()
    assert t.reduce("max").flatArray()[0] == x.max()
    assert (t * T.fill(2.0, shape)).reduce("max").flatArray()[0] == (x * np.float32(2.0)).max()
    small = T.random([10, 12], seed=4) + T.fill(0.5, [10, 12])
    want = np.prod((ref.random_buffer(120, 4) + np.float32(0.5)).astype(np.float64))
    assert abs(float(small.product().flatArray()[0]) - want) <= 1e-5 * abs(want)
    # a view with padding inside the fold: translate pulls in the padding value (here 2.0) on the shifted border
    p = T.random([16, 24], seed=6, padding=2.0).doCache()
    v = p.translate([3, -5])
    want_v = ref.Tensor.random([16, 24], seed=6, padding=2.0).translate([3, -5]).flat_array()
    got = v.sum().flatArray()[0]
    assert abs(float(got) - float(want_v.astype(np.float64).sum())) <= 1e-5 * float(want_v.astype(np.float64).sum())
    assert v.reduce("max").flatArray()[0] == 2.0 and v.reduce("min").flatArray()[0] == want_v.min()
    # permuted operand (non-flat index decode) on exactly summable data
    e = dataset_e(T, [24, 32, 8]).doCache()
    assert e.permute([2, 0, 1]).sum().flatArray()[0] == np.float32(dataset_e_np(24 * 32 * 8).astype(np.int64).sum())
    with pytest.raises(ValueError):
        t.reduce("-")


def axis_sum(T, x, axis):
    parts = x.split(axis)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


@pytest.mark.parametrize("shape,axis", [((512, 1024), 0), ((512, 1024), 1), ((256, 8, 12), 0), ((33, 70), 1), ((6, 9), 0), ((100, 64), 0),
                                        ((4099, 260), 0), ((3000, 7), 0)])
def test_c3_axis_sums(cuda, shape, axis):
    T = cuda.Tensor
    n = int(np.prod(shape))
    e = dataset_e(T, list(shape)).doCache()
    k = axis_sum(T, e, axis).compile()
    assert k.info.kind == (1 if shape[axis] >= 8 else 0)  # re-rolled into a real reduction
    got = axis_sum(T, e, axis).flatArray()
    want = dataset_e_np(n).reshape(shape).astype(np.int64).sum(axis=axis).astype(np.float32).reshape(-1)
    assert np.array_equal(got, want)  # bit-exact on exactly summable data
    u = T.random(list(shape), seed=5).doCache()
    got_u = axis_sum(T, u, axis).flatArray().astype(np.float64)
    x = ref.random_buffer(n, 5).reshape(shape)
    truth = x.astype(np.float64).sum(axis=axis).reshape(-1)
    left_fold = np.add.accumulate(np.moveaxis(x, axis, 0), axis=0, dtype=np.float32)[-1].reshape(-1)  # reference order
    assert np.abs(got_u - truth).max() <= 1e-5 * np.abs(truth).max()
    assert np.abs(got_u - left_fold.astype(np.float64)).max() <= 2e-5 * np.abs(truth).max()


def test_axis_folds_with_every_monoid(cuda):
    """t.split(axis).reduce(max | min | *) (MonoidPrograms, T:308-311): min / max are order-independent -> bit-exact vs the
    oracle's unrolled chain, NaN handling included; products within 1e-5"""
    T = cuda.Tensor

    def chain(parts, f):
        acc = parts[0]
        for p in parts[1:]:
            acc = f(acc, p)
        return acc

    def data(T, shape):
        return T.random(shape, seed=21) * T.fill(2.0, shape) - T.fill(1.0, shape)

    for shape in ([64, 33], [9, 40, 12], [300, 8]):
        for axis in range(len(shape)):
            if shape[axis] < 8:
                continue
            for f in ("max", "min"):
                def build(T):
                    return chain(data(T, shape).split(axis), getattr(T, f))
                got, want = build(T).flatArray(), build(ref.Tensor).flat_array()
                assert build(T).compile().info.kind == 1
                assert np.array_equal(bits(got), bits(want)), (shape, axis, f)

            def prod(T):
                x = data(T, shape) * T.fill(0.25, shape) + T.fill(1.0, shape)  # factors in [0.75, 1.25]
                return chain(x.split(axis), lambda a, b: a * b)
            got, want = prod(T).flatArray().astype(np.float64), prod(ref.Tensor).flat_array().astype(np.float64)
            assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # other shapes of the same fold: pairwise (parallel reduce) and right fold, exact data so that the order cannot matter
    def pairwise(parts, f):
        while len(parts) > 1:
            parts = [f(parts[i], parts[i + 1]) if i + 1 < len(parts) else parts[i] for i in range(0, len(parts), 2)]
        return parts[0]

    def right(parts, f):
        acc = parts[-1]
        for q in reversed(parts[:-1]):
            acc = f(q, acc)
        return acc

    for red in (pairwise, right):
        for axis in (0, 1):
            def build(T):
                return T.fill(0.5, [37 if axis == 0 else 50]) * red(dataset_e(T, [50, 37]).split(axis), lambda a, b: a + b)
            got, want = build(T).flatArray(), build(ref.Tensor).flat_array()
            assert build(T).compile().info.kind == 1 and np.array_equal(bits(got), bits(want)), (red.__name__, axis)
    # NaN semantics of fmin / fmax (K:271-325 -> OpenCL fmin/fmax): NaNs are ignored unless every term is NaN
    x = np.arange(24 * 16, dtype=np.float32).reshape(24, 16) % 7 - 3
    x[3, :] = np.nan      # one whole reduction column of the axis-1 fold ...
    x[:, 5] = np.nan      # ... and one of the axis-0 fold; scattered ones elsewhere
    x[7, 2] = np.nan
    for axis in (0, 1):
        for f in ("max", "min"):
            got = chain(T(x).split(axis), getattr(T, f)).flatArray()
            want = chain(ref.Tensor(x).split(axis), getattr(ref.Tensor, f)).flat_array()
            assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), (axis, f)
            assert np.isnan(got).sum() == 1
    # the softmax idiom, rows of 512: max-shift, exponentials, row sums, quotient -- three reductions / epilogues, no unrolled kernels
    def softmax(T):
        shape = [96, 512]
        z = data(T, shape) * T.fill(8.0, shape)
        m = chain(z.split(1), T.max)
        e = T.exp(z - m.broadcast(shape))
        s = chain(e.split(1), lambda a, b: a + b)
        return e / s.broadcast(shape)
    got = softmax(T).flatArray().astype(np.float64).reshape(96, 512)
    want = softmax(ref.Tensor).flat_array().astype(np.float64).reshape(96, 512)
    assert np.abs(got.sum(axis=1) - 1.0).max() < 1e-5 and np.abs(got - want).max() <= 1e-6


# ---- C4: views ------------------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n", [64, 20])
def test_c4_views_bit_exact(cuda, n):
    def permute_translate(T):
        t = T.random([n, n, n], seed=7)
        return t.permute([2, 0, 1]).translate([3, -5, 7])

    def trailing(T):
        return T.random([n, n], seed=8).broadcast([n, n, n])

    def leading(T):
        return T.random([n, n], seed=8).reshape([1, n, n]).broadcast([n, n, n])

    def split_join(T):
        return T.join(T.random([n, n, n], seed=7).split(1))

    def split_join_mid(T):
        return T.join(T.random([n, n, n], seed=7).split(1), 1)

    def fused_views(T):  # views feeding arithmetic in one kernel
        t = T.random([n, n, n], seed=7)
        m = T.random([n, n], seed=8)
        return t.transpose() * m.broadcast([n, n, n]) + t.translate([1, 0, -2])

    for build in (permute_translate, trailing, leading, split_join, split_join_mid):
        got = build(cuda.Tensor).flatArray()
        want = build(ref.Tensor).flat_array()
        assert np.array_equal(bits(got), bits(want)), build.__name__
    got = fused_views(cuda.Tensor).flatArray()
    assert min_ulp(got, oracle_bracket(fused_views)).max() <= 1


def test_c4_roundtrip_is_identity(cuda):
    T = cuda.Tensor
    t = T.random([24, 16, 40], seed=7).doCache()
    base = t.flatArray()
    assert np.array_equal(bits(T.join(t.split(1), 1).flatArray()), bits(base))
    assert np.array_equal(bits(t.permute([2, 0, 1]).permute([1, 2, 0]).flatArray()), bits(base))
    assert np.array_equal(bits(t.translate([1, -2, 3]).translate([-1, 2, -3]).flatArray()[:0]), bits(base)[:0])


def test_scale_non_integer_coefficients(cuda):
    """Tensors.scala:950-965 + the DecimalFormat / (int) truncation quirks of OpenCLKernelBuilder.scala:14-32,386"""
    for src, dst in (([6, 9], [4, 3]), ([5, 5], [7, 8]), ([3, 4], [9, 16])):

        def build(T):
            return T.random(src, seed=21).scale(dst)

        got = build(cuda.Tensor).flatArray()
        want = build(ref.Tensor).flat_array()
        assert np.array_equal(bits(got), bits(want)), (src, dst)


# ---- C5: matmul expressed as split / broadcast / sum ---------------------------------------------------------------------------------


def matmul2(T, a, b):  # benchmarks.scala:188-191
    i, j = a.shape
    _, k = b.shape
    product = a.broadcast([i, j, k]) * b.reshape([1, j, k]).broadcast([i, j, k])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def left_fold_matmul(a, b):
    """restated reference arithmetic: C[i,k] = fp32 left fold over t of A[i,t]*B[t,k]"""
    acc = (a[:, 0:1] * b[0:1, :]).astype(np.float32)
    for t in range(1, a.shape[1]):
        acc = (acc + (a[:, t : t + 1] * b[t : t + 1, :]).astype(np.float32)).astype(np.float32)
    return acc


@pytest.mark.parametrize("m,k,n", [(128, 256, 128), (256, 512, 384), (48, 40, 24), (5, 9, 7)])
def test_c5_matmul_pattern(cuda, m, k, n):
    T = cuda.Tensor
    rng = np.random.default_rng(9)
    a_e = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b_e = rng.integers(-4, 5, (k, n)).astype(np.float32)
    got = matmul2(T, T(a_e), T(b_e)).flatArray().reshape(m, n)
    assert np.array_equal(got, left_fold_matmul(a_e, b_e))  # dataset E: bit-exact
    assert np.array_equal(got, (a_e.astype(np.int64) @ b_e.astype(np.int64)).astype(np.float32))
    # randomNormal yields +-inf / nan where the hash hits u1 == 0 (log(0), Tensors.scala:415-417): keep the data finite
    a_n = np.nan_to_num(ref.random_normal_buffer(m * k, 9), nan=0.0, posinf=0.0, neginf=0.0).reshape(m, k)
    b_n = np.nan_to_num(ref.random_normal_buffer(k * n, 10), nan=0.0, posinf=0.0, neginf=0.0).reshape(k, n)
    got = matmul2(T, T(a_n), T(b_n)).flatArray().reshape(m, n).astype(np.float64)
    want = left_fold_matmul(a_n, b_n).astype(np.float64)
    scale = np.abs(a_n.astype(np.float64)) @ np.abs(b_n.astype(np.float64))
    assert (np.abs(got - want) / scale).max() <= 1e-5
    truth = a_n.astype(np.float64) @ b_n.astype(np.float64)
    assert (np.abs(got - truth) / scale).max() <= 1e-5


def test_c5_never_materialises_the_product(cuda):
    """matmul2's i*j*k intermediate (SURVEY finding 2) must not be allocated: the pattern is composed into the reduction"""
    T = cuda.Tensor
    m = k = n = 256
    a, b = T.randomNormal([m, k], seed=9).doCache(), T.randomNormal([k, n], seed=10).doCache()
    cuda.synchronize()
    before = cuda.stats()
    kern = matmul2(T, a, b).compile()
    assert kern.info.kind in (1, 2)
    assert kern.info.n_args == 2
    out = matmul2(T, a, b).flatArray()
    after = cuda.stats()
    assert out.size == m * n
    assert after["bytes_in_use"] == before["bytes_in_use"]


def _convolute(T, inp, weight, bias):
    """benchmarks.scala:463-556 (the 8f-2 workload): split + translate + broadcast + join with padding"""
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


@pytest.mark.parametrize("batch,size,depth,filters,ks", [(4, 16, 8, 8, 3), (3, 9, 5, 6, 3), (2, 8, 16, 4, 1), (2, 12, 3, 7, 5), (8, 24, 32, 32, 3),
                                                         (5, 20, 24, 40, 3)])
def test_convolution_nested_reduction(cuda, batch, size, depth, filters, ks):
    """the re-rolled (kernel row x kernel column x channel) reduction with its bias epilogue against a direct numpy convolution
    on exactly representable data (bit-exact in any order) and against the oracle's literal evaluation of the same graph"""
    T = cuda.Tensor
    rng = np.random.default_rng(batch + size + depth)
    inp = rng.integers(-3, 4, (batch, size, size, depth)).astype(np.float32)
    w = rng.integers(-3, 4, (ks, ks, depth, filters)).astype(np.float32)
    b = rng.integers(-8, 9, (filters,)).astype(np.float32) / np.float32(2.0)
    e = _convolute(T, T(inp), T(w), T(b))
    macs = batch * size * size * filters * ks * ks * depth
    implicit_gemm = macs >= 2**25 and filters >= 32 and ks * ks * depth >= 32  # gathered panels -> tcgen05 (an implicit im2col)
    assert e.compile().info.kind == (2 if implicit_gemm else 1 if ks * ks * depth >= 8 else 0)
    got = e.flatArray().reshape(batch, size, size, filters)
    r = ks // 2
    padded = np.zeros((batch, size + 2 * r, size + 2 * r, depth), np.float64)
    padded[:, r : r + size, r : r + size, :] = inp
    want = np.zeros((batch, size, size, filters), np.float64) + b
    for ky in range(ks):
        for kx in range(ks):
            # out[y, x] += in[y - oy, x - ox] * w[ky, kx] with oy = ky - r (translate semantics, Tensors.scala:970-976)
            oy, ox = ky - r, kx - r
            window = padded[:, r - oy : r - oy + size, r - ox : r - ox + size, :]
            want += np.einsum("bhwd,df->bhwf", window, w[ky, kx].astype(np.float64))
    assert np.array_equal(got.astype(np.float64), want)
    if batch * size * size * filters * ks * ks * depth <= 200000:
        R = ref.Tensor
        lit = _convolute(R, R(inp), R(w), R(b)).flat_array()
        assert np.array_equal(got.reshape(-1).view(np.uint32), lit.view(np.uint32))


@pytest.mark.parametrize("batch,size,depth,filters,ks", [(128, 32, 8, 8, 3), (32, 32, 8, 8, 3), (128, 32, 8, 8, 1), (17, 31, 8, 12, 3), (20, 32, 6, 6, 3),
                                                         (9, 32, 16, 16, 1), (6, 40, 4, 32, 3), (8, 64, 16, 16, 3), (5, 40, 28, 24, 3)])
def test_convolution_small_n_on_warp_mmas(cuda, batch, size, depth, filters, ks):
    """the reference's own benchmark sizes (benchmarks.scala:612-630) and ragged relatives: many pixels, few filters, short K. One generated
    kernel keeps the weights as TF32 hi / lo fragments in registers and streams the pixels through warp-level MMAs (3xTF32): exact on
    exactly representable data (rows not a multiple of 16, filters not a multiple of 8, K padded to 8 included), <= 1e-5 of |in| |w| on
    normal data (reductions and matmul), and the lo terms are really there (hi-only would be ~1e-3)"""
    T = cuda.Tensor
    rng = np.random.default_rng(batch + size + depth)
    r = ks // 2

    def direct(inp, w, b):
        padded = np.zeros((batch, size + 2 * r, size + 2 * r, depth), np.float64)
        padded[:, r : r + size, r : r + size, :] = inp
        want = np.zeros((batch, size, size, filters), np.float64) + b
        mag = np.zeros((batch, size, size, filters), np.float64) + np.abs(b)
        for ky in range(ks):
            for kx in range(ks):
                oy, ox = ky - r, kx - r
                window = padded[:, r - oy : r - oy + size, r - ox : r - ox + size, :]
                want += np.einsum("bhwd,df->bhwf", window, w[ky, kx].astype(np.float64))
                mag += np.einsum("bhwd,df->bhwf", np.abs(window), np.abs(w[ky, kx]).astype(np.float64))
        return want, mag

    inp = rng.integers(-3, 4, (batch, size, size, depth)).astype(np.float32)
    w = rng.integers(-3, 4, (ks, ks, depth, filters)).astype(np.float32)
    b = rng.integers(-8, 9, (filters,)).astype(np.float32) / np.float32(2.0)
    e = _convolute(T, T(inp), T(w), T(b))
    kern = e.compile()
    assert kern.info.kind == 1 and "small-N contraction" in kern.source and "cc_mma_tf32_16x8x8" in kern.source, kern.source[:200]
    got = e.flatArray().reshape(batch, size, size, filters)
    assert np.array_equal(got.astype(np.float64), direct(inp, w, b)[0])
    inp = rng.standard_normal((batch, size, size, depth)).astype(np.float32)
    w = (rng.standard_normal((ks, ks, depth, filters)) * (1.0 + 2.0**-12)).astype(np.float32)
    b = rng.standard_normal((filters,)).astype(np.float32)
    got = _convolute(T, T(inp), T(w), T(b)).flatArray().reshape(batch, size, size, filters)
    want, mag = direct(inp, w, b)
    assert (np.abs(got - want) / mag).max() <= 2e-6  # (the bar is 1e-5; fp32 FMA chains of this length sit at ~1e-7)


@pytest.mark.parametrize("m,k,n", [(4096, 20, 12), (5000, 33, 8), (65536, 32, 32), (4100, 256, 6), (8192, 64, 16), (4097, 9, 30)])
def test_small_n_contraction_over_views(cuda, m, k, n):
    """skinny products written as split / broadcast / sum (benchmarks.scala:188-191) whose operands are views: A stored transposed, B read
    through a translation with a non-zero padding, an epilogue around the sum. Tall M, N <= 32, short K: the small-N kernel on warp-level
    MMAs (K not a multiple of 8, rows not a multiple of 16, N not a multiple of 8 included), exact on exactly representable data"""
    T = cuda.Tensor
    rng = np.random.default_rng(m + k + n)
    at = rng.integers(-4, 5, (k, m)).astype(np.float32)       # A^T in memory
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    bias = rng.integers(-6, 7, (n,)).astype(np.float32) / np.float32(4.0)
    A = T(at).permute([1, 0])                                  # [m, k] view
    Bs = T(b, padding=2.5).translate([1, 0])                   # rows shifted down by one, first row = padding 2.5
    prod = A.broadcast([m, k, n]) * Bs.nonInline().reshape([1, k, n]).broadcast([m, k, n])
    acc = axis_sum(T, prod, 1) * T.fill(0.5, [m, n]) + T(bias).reshape([1, n]).broadcast([m, n])
    kern = acc.compile()
    assert kern.info.kind == 1 and "small-N contraction" in kern.source, kern.source[:300]
    bs_host = np.full_like(b, 2.5)
    bs_host[1:] = b[:-1]
    want = ((at.T.astype(np.float64) @ bs_host.astype(np.float64)) * 0.5 + bias).astype(np.float32)
    assert np.array_equal(acc.flatArray().reshape(m, n), want)
    # the plain operands, the pattern as the reference writes it
    a = np.ascontiguousarray(at.T)
    plain = axis_sum(T, T(a).broadcast([m, k, n]) * T(b).reshape([1, k, n]).broadcast([m, k, n]), 1)
    assert "small-N contraction" in plain.compile().source
    assert np.array_equal(plain.flatArray().reshape(m, n), (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("m,k,f,fold", [(20000, 16, 8, "sum"), (19000, 12, 16, "max"), (40000, 20, 4, "sum")])
def test_tile_owner_reductions_that_are_not_products(cuda, m, k, f, fold):
    """distances of many points to a few centres, written as broadcast / split / fold like the matmul pattern but with |x - c| (or its max)
    instead of x * c: not a contraction, so the re-rolled reduction's tile owner runs — a thread owns all f outputs of a point, the point's
    coordinates are fetched as vectors along the reduction and shared by the f centres. Exact on exactly representable data."""
    T = cuda.Tensor
    rng = np.random.default_rng(m + k + f)
    x = rng.integers(-9, 10, (m, k)).astype(np.float32)
    c = rng.integers(-9, 10, (k, f)).astype(np.float32)
    d = T.abs(T(x).broadcast([m, k, f]) - T(c).reshape([1, k, f]).broadcast([m, k, f]))
    parts = d.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p if fold == "sum" else T.max(acc, p)
    kern = acc.compile()
    assert kern.info.kind == 1 and "tile owner" in kern.source, kern.source[:300]
    diff = np.abs(x[:, :, None].astype(np.float64) - c[None, :, :].astype(np.float64))
    want = diff.sum(axis=1) if fold == "sum" else diff.max(axis=1)
    assert np.array_equal(acc.flatArray().reshape(m, f).astype(np.float64), want)


@pytest.mark.parametrize("rows", [64, 4096])  # one CTA covers all of T / partials + second stage (epilogue applied there)
def test_epilogue_around_an_axis_sum(cuda, rows):
    T = cuda.Tensor
    x = dataset_e(T, [rows, 512]).doCache()
    b = T.random([512], seed=2).doCache()
    e = T.tanh(axis_sum(T, x, 0) * T.fill(1.0 / 64.0, [512]) + b)
    assert e.compile().info.kind == 1
    got = e.flatArray()
    cols = dataset_e_np(rows * 512).reshape(rows, 512).astype(np.int64).sum(axis=0).astype(np.float32)
    want = np.tanh((cols * np.float32(1.0 / 64.0) + ref.random_buffer(512, 2)).astype(np.float64))
    assert np.abs(got - want).max() <= 4e-7  # tanh within 2 ulp of values in [-1, 1]


def test_general_contraction_over_views(cuda):
    """a matmul whose operands are views (A stored transposed, B read through a translate with padding) is still a GEMM over gathered
    operands: generated kernels write the K-major TF32 panels from the affine maps and the tcgen05 pipeline runs on them"""
    T = cuda.Tensor
    m, k, n = 384, 320, 288
    rng = np.random.default_rng(7)
    at = rng.integers(-4, 5, (k, m)).astype(np.float32)       # A^T in memory
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    A = T(at).permute([1, 0])                                  # [m, k] view
    Bs = T(b).translate([2, 0])                                # rows shifted down by 2, first two rows = padding 0
    # written the way benchmarks.scala:188-191 writes it, over the views
    a3 = A.broadcast([m, k, n])
    b3 = Bs.nonInline().reshape([1, k, n]).broadcast([m, k, n])
    acc = axis_sum(T, a3 * b3, 1)
    bs_host = np.zeros_like(b)
    bs_host[2:] = b[:-2]
    want = (at.T.astype(np.float64) @ bs_host.astype(np.float64)).astype(np.float32)
    assert np.array_equal(acc.flatArray().reshape(m, n), want)
    # the transposed operand alone (no materialisation of A): general contraction
    acc2 = axis_sum(T, A.broadcast([m, k, n]) * T(b).reshape([1, k, n]).broadcast([m, k, n]), 1)
    kern = acc2.compile()
    assert kern.info.kind == 2 and "general contraction" in kern.source, kern.source[:300]
    assert np.array_equal(acc2.flatArray().reshape(m, n), (at.T.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
