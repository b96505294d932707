"""GPU parity at BASELINE.json's FULL sizes (C2 2^28 elements, C3 16384^2, C4 512^3; C5 8192^3 lives in test_gemm.py):
against the C port of the oracle where it finishes in seconds on the host cores, and through size-independent properties
(exact integer sums, round trips, checksums) elsewhere."""
import numpy as np
import pytest

from oracle import build as ob
from oracle import reference as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    from compute.scala_b200 import cuda as c

    c.init()
    yield c
    c.synchronize()


def host_random(n, seed):
    out = np.empty(n, np.float32)
    ob.load("strict").oracle_random(out.ctypes.data, n, seed & 0xFFFFFFFF)
    return out


def test_c2_chain_2_pow_28(cuda):
    T = cuda.Tensor
    shape = [16384, 16384]
    n = 1 << 28
    a, b, c = (T.random(shape, seed=s).doCache() for s in (1, 2, 3))
    got = (T.tanh(T.log(T.exp(a * b + c) + a) * b) + c).flatArray()
    ha, hb, hc = (host_random(n, s) for s in (1, 2, 3))
    assert np.array_equal(a.flatArray().view(np.uint32), ha.view(np.uint32))  # inputs: bit-exact with the reference's random
    worst = 0.0
    for variant in ("strict", "fma"):
        want = np.empty(n, np.float32)
        ob.load(variant).oracle_c2(ha.ctypes.data, hb.ctypes.data, hc.ctypes.data, want.ctypes.data, n)
        worst = max(worst, float(np.abs(got - want).max()))
    assert worst <= 5e-6, worst  # absolute: the chain is ill-conditioned in ulp terms near log(1+eps)
    # the 2-ulp-per-op forward error bound (see test_parity_configs.c2_truth_and_bound) on a strided sample of 2^22 elements
    idx = np.arange(0, n, 64)
    x, y, z = (v[idx].astype(np.float64) for v in (ha, hb, hc))
    u24, F = 2.0**-24, 2.0 * 2.0**-23
    ab = x * y
    t = ab + z
    e = (np.abs(ab) + np.abs(t)) * u24
    u = np.exp(t)
    e = u * e + u * F
    s_ = u + x
    e = e + np.abs(s_) * u24
    v = np.log(s_)
    e = e / s_ + np.abs(v) * F
    p = v * y
    e = np.abs(y) * e + np.abs(p) * u24
    w = np.tanh(p)
    e = (1 - w * w) * e + np.abs(w) * F
    out = w + z
    e = 1.05 * (e + np.abs(out) * u24)
    assert (np.abs(got[idx].astype(np.float64) - out) <= e).all()


def test_c3_sums_16384_squared(cuda):
    T = cuda.Tensor
    rows = cols = 16384
    n = rows * cols
    r = T.random([rows, cols], seed=5)
    e = ((r * T.fill(9.0, [rows, cols])) - (r * T.fill(9.0, [rows, cols])) % T.fill(1.0, [rows, cols]) - T.fill(4.0, [rows, cols])).doCache()
    he = (np.floor(host_random(n, 5) * np.float32(9.0)) - np.float32(4.0)).astype(np.float32)
    hi = he.astype(np.int64).reshape(rows, cols)
    assert np.abs(np.cumsum(he[: 1 << 20].astype(np.int64))).max() < 2**24  # partial sums stay exactly representable
    assert e.sum().flatArray()[0] == np.float32(hi.sum())  # bit-exact, any order
    # the same sum with the dataset's closure fused into the fold kernel (nothing materialised), and the other monoids
    r2 = T.random([rows, cols], seed=5)
    inline_e = (r2 * T.fill(9.0, [rows, cols])) - (r2 * T.fill(9.0, [rows, cols])) % T.fill(1.0, [rows, cols]) - T.fill(4.0, [rows, cols])
    fused = inline_e.sum()
    assert fused.compile().info.kind == 4
    assert fused.flatArray()[0] == np.float32(hi.sum())
    assert e.reduce("max").flatArray()[0] == he.max() and e.reduce("min").flatArray()[0] == he.min()  # (r * 9 can round up to 9.0)

    def axis_sum(x, axis):
        parts = x.split(axis)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    assert np.array_equal(axis_sum(e, 0).flatArray(), hi.sum(axis=0).astype(np.float32))
    assert np.array_equal(axis_sum(e, 1).flatArray(), hi.sum(axis=1).astype(np.float32))
    # dataset U: uniform [0,1): <= 1e-5 relative to the fp64 truth; the reference's own 16-lane order is reported beside it
    u = r.doCache()
    hu = host_random(n, 5)
    L = ob.load("strict")
    truth = L.oracle_sum_fp64(hu.ctypes.data, n)
    ref_order = float(L.oracle_sum_cpu_order(hu.ctypes.data, n))
    got = float(u.sum().flatArray()[0])
    assert abs(got - truth) <= 1e-5 * truth, (got, truth)
    print(f"full sum: cuda rel err {abs(got - truth) / truth:.2e}, reference CPU order rel err {abs(ref_order - truth) / truth:.2e}")
    assert abs(got - truth) <= abs(ref_order - truth) + 1e-7 * truth  # no further from the truth than the reference's own order
    cs = axis_sum(u, 0).flatArray().astype(np.float64)
    rs = axis_sum(u, 1).flatArray().astype(np.float64)
    h2 = hu.reshape(rows, cols).astype(np.float64)
    assert np.abs(cs - h2.sum(axis=0)).max() <= 1e-5 * h2.sum(axis=0).max()
    assert np.abs(rs - h2.sum(axis=1)).max() <= 1e-5 * h2.sum(axis=1).max()
    assert abs(cs.sum() - truth) <= 1e-5 * truth and abs(rs.sum() - truth) <= 1e-5 * truth  # checksum of checksums


def test_c4_views_512_cubed(cuda):
    T = cuda.Tensor
    d = 512
    t = T.random([d, d, d], seed=7).doCache()
    m = T.random([d, d], seed=8).doCache()
    ht = host_random(d**3, 7).reshape(d, d, d)
    hm = host_random(d * d, 8).reshape(d, d)
    bits = lambda x: np.ascontiguousarray(x, dtype=np.float32).reshape(-1).view(np.uint32)
    # (i) permute(2,0,1) then translate(3,-5,7): out[g0,g1,g2] = T[g1+5, g2-7, g0-3] or padding 0
    got = t.permute([2, 0, 1]).translate([3, -5, 7]).flatArray()
    want = np.empty(d**3, np.float32)
    mat = np.asarray([[0, 1, 0, 5], [0, 0, 1, -7], [1, 0, 0, -3]], np.int64)
    shp = np.asarray([d, d, d], np.int64)
    ob.load("strict").oracle_affine_gather_3d(ht.ctypes.data, shp.ctypes.data, 3, mat.ctypes.data, shp.ctypes.data, 0.0, want.ctypes.data)
    assert np.array_equal(bits(got), bits(want))
    # (ii) trailing and leading broadcast
    assert np.array_equal(bits(m.broadcast([d, d, d]).flatArray()), bits(np.broadcast_to(hm[:, :, None], (d, d, d))))
    assert np.array_equal(bits(m.reshape([1, d, d]).broadcast([d, d, d]).flatArray()), bits(np.broadcast_to(hm[None], (d, d, d))))
    # (iii) split / join: join(split(1)) moves dim 1 last; join(split(1), 1) is the identity
    assert np.array_equal(bits(T.join(t.split(1)).flatArray()), bits(ht.transpose(0, 2, 1)))
    assert np.array_equal(bits(T.join(t.split(1), 1).flatArray()), bits(ht))
    assert np.array_equal(bits(t.permute([2, 0, 1]).permute([1, 2, 0]).flatArray()), bits(ht))


def test_c5_matmul_pattern_8192(cuda):
    """C5 at BASELINE size THROUGH the Tensor API, written as the reference writes it (benchmarks.scala:188-191): broadcast both
    operands to [i, j, k], multiply, split(1), fold with +. Dataset E (integers in {-4..5}: every summation order is exact), checked on
    sampled rows and through the checksum identity sum(C) = colsum(A) . rowsum(B)."""
    T = cuda.Tensor
    n = 8192
    nine, four, one = T.fill(9.0, [n, n]), T.fill(4.0, [n, n]), T.fill(1.0, [n, n])

    def dataset_e(seed):
        r = T.random([n, n], seed=seed) * nine
        return ((r - r % one) - four).doCache()

    A, B = dataset_e(9), dataset_e(10)
    product = A.broadcast([n, n, n]) * B.reshape([1, n, n]).broadcast([n, n, n])  # lazy: 2 TiB if it were ever materialised
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    k = acc.compile()
    assert k.info.kind == 2 and k.info.n_args == 2 and k.info.flops == 2 * n**3
    c = acc.flatArray().reshape(n, n)
    a, b = A.flatArray().reshape(n, n), B.flatArray().reshape(n, n)
    rows = np.r_[0:4, 255:258, 4095:4098, 8188:8192]
    assert np.array_equal(c[rows].astype(np.float64), a[rows].astype(np.float64) @ b.astype(np.float64))
    assert float(c.astype(np.float64).sum()) == float(a.astype(np.float64).sum(axis=0) @ b.astype(np.float64).sum(axis=1))


def test_c5_matmul_pattern_8192_dataset_n(cuda):
    """C5 at BASELINE size on dataset N (A, B = randomNormal(seed 9, 10)) THROUGH the split / broadcast / sum pattern
    (benchmarks.scala:188-191), i.e. on the CTA-pair kernel with non-zero lo panels. Bar: max|D| / (|A| . |B|) <= 1e-5 against the
    reference's fp32 left fold (oracle_matmul_left_fold restates the generated kernel's `acc = acc + a*b` chain) and against fp64,
    on sampled rows spread over both CTAs of a pair, several tiles and the matrix edges; plus the checksum identity in fp64."""
    T = cuda.Tensor
    n = 8192
    # randomNormal's singular pair (hash(61) = 0 => elements 2 * (61 ^ seed), +1 are (+inf, NaN), in the reference too) is zeroed on the host
    def dataset_n(seed):
        h = T.randomNormal([n, n], seed=seed).flatArray()
        assert np.allclose(h[:4096], ref.random_normal_buffer(4096, seed), rtol=1e-4, atol=1e-6, equal_nan=True)  # the reference's stream
        assert (~np.isfinite(h)).sum() <= 2
        h = np.nan_to_num(h, nan=0.0, posinf=0.0, neginf=0.0).reshape(n, n)
        return T(h).doCache(), h

    (A, a), (B, b) = dataset_n(9), dataset_n(10)
    product = A.broadcast([n, n, n]) * B.reshape([1, n, n]).broadcast([n, n, n])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    kern = acc.compile()
    assert kern.info.kind == 2 and kern.info.flops == 2 * n**3
    c = acc.flatArray().reshape(n, n)
    rows = np.r_[0:2, 127:130, 255:258, 4000:4002, 6143:6146, 8190:8192]
    ar = np.ascontiguousarray(a[rows])
    a64, b64 = ar.astype(np.float64), b.astype(np.float64)
    truth = a64 @ b64
    scale = np.abs(a64) @ np.abs(b64)
    got = c[rows].astype(np.float64)
    err64 = float((np.abs(got - truth) / scale).max())
    lf = np.empty((len(rows), n), np.float32)
    ob.load("strict").oracle_matmul_left_fold(ar.ctypes.data, b.ctypes.data, lf.ctypes.data, len(rows), n, n)
    err_lf = float((np.abs(got - lf.astype(np.float64)) / scale).max())
    ref_err = float((np.abs(lf.astype(np.float64) - truth) / scale).max())
    print(f"C5 dataset N: cuda vs fp64 {err64:.2e}, cuda vs reference left fold {err_lf:.2e}, reference left fold vs fp64 {ref_err:.2e}")
    assert err_lf <= 1e-5 and err64 <= 1e-5  # the north star's bar
    # (measured 4.9e-6 at K = 8192: the tensor core accumulates its 3 * K / 8 partial products per output in fp32 with truncation, so the
    # error grows linearly in K where the reference's round-to-nearest left fold grows like sqrt(K); single-pass TF32 sits near 2e-4)
    assert err64 <= 6e-6
    # checksum over the whole result (every tile of every CTA): sum(C) = colsum(A) . rowsum(B), both sides in fp64
    total = float(c.astype(np.float64).sum())
    want_total = float(a.astype(np.float64).sum(axis=0) @ b.astype(np.float64).sum(axis=1))
    bound = 1e-5 * float(np.abs(a).astype(np.float64).sum(axis=0) @ np.abs(b).astype(np.float64).sum(axis=1)) / n  # errors average out
    assert abs(total - want_total) <= bound, (total, want_total, bound)
