"""GPU parity for the tcgen05 3xTF32 contraction (C5): called directly through cc_matmul_3xtf32 and through the
split / broadcast / sum pattern, against the restated reference arithmetic (fp32 left fold)."""
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


def run_matmul(cuda, a, b):
    m, k = a.shape
    _, n = b.shape
    A, B, C = cuda.Buffer.from_host(a), cuda.Buffer.from_host(b), cuda.Buffer.alloc(m * n)
    cuda.matmul_3xtf32(A, B, C, m, n, k)
    out = C.to_host(m * n).reshape(m, n)
    for x in (A, B, C):
        x.release()
    return out


def finite_normal(n, seed):
    return np.nan_to_num(ref.random_normal_buffer(n, seed), nan=0.0, posinf=0.0, neginf=0.0)


@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (128, 256, 64), (256, 512, 96), (384, 256, 1024), (1024, 1024, 1024)])
def test_exact_on_small_integers(cuda, m, n, k):
    rng = np.random.default_rng(m + n + k)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    got = run_matmul(cuda, a, b)
    want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)  # exact: |sums| < 2^24
    assert np.array_equal(got, want)


# ragged shapes: M / N edges are TMA out-of-bounds zero fill in + predicated stores out, K is zero-padded in the workspace;
# the tile width (256 / 128 / 64) is picked per problem, so these also cover every instantiation of the pipeline; the last four
# are large enough for 256-wide tiles and run on CTA pairs (tcgen05.mma.cta_group::2), with ragged M / N / K and an odd tile count
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (5, 7, 9), (127, 255, 31), (129, 257, 33), (200, 100, 50), (1000, 1000, 1000), (333, 64, 70),
                                   (2048, 96, 40), (77, 1030, 129), (4096, 32, 32), (130, 66, 2051), (19000, 130, 64), (20000, 520, 40), (19072, 256, 32),
                                   (37888, 512, 256)])
def test_exact_on_small_integers_any_shape(cuda, m, n, k):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    got = run_matmul(cuda, a, b)
    want = (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("config", [1024, 512, 256, 128, 64])
@pytest.mark.parametrize("m,n,k", [(300, 260, 40), (129, 1, 1), (1000, 513, 200), (2048, 2048, 96), (1030, 136, 520), (4096, 4096, 128)])
def test_every_tile_configuration_on_the_same_shapes(cuda, monkeypatch, config, m, n, k):
    """1024 = 256x256 tiles on CTA pairs with A read as the original fp32 matrix and split through tensor memory inside the kernel (needs
    K % 4 == 0, else the request falls back to the picker's choice); 512 = 256x256 tiles on CTA pairs (tcgen05.mma.cta_group::2, 3 stages),
    256 / 128 / 64 = one CTA per 128 x that many columns"""
    monkeypatch.setenv("CC_GEMM_FORCE_CONFIG", str(config))
    rng = np.random.default_rng(m + n + k + config)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    got = run_matmul(cuda, a, b)
    assert np.array_equal(got, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))


def test_ragged_edges_do_not_write_outside_the_result(cuda):
    m, n, k = 130, 70, 33
    rng = np.random.default_rng(1)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    guard = 4096
    A, B = cuda.Buffer.from_host(a), cuda.Buffer.from_host(b)
    C = cuda.Buffer.from_host(np.full(m * n + guard, -777.0, np.float32))
    cuda.matmul_3xtf32(A, B, C, m, n, k)
    out = C.to_host(m * n + guard)
    assert np.array_equal(out[: m * n].reshape(m, n), (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
    assert (out[m * n:] == -777.0).all()
    for x in (A, B, C):
        x.release()


def test_b_panels_are_cached_while_b_is_unchanged(cuda):
    """the hi / lo split of an unchanged B (weights, the replicated operand of the sharded matmul) runs once; any write to B
    (here an upload into the same buffer) invalidates it"""
    m, n, k = 256, 192, 100
    rng = np.random.default_rng(5)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b1 = rng.integers(-4, 5, (k, n)).astype(np.float32)
    b2 = rng.integers(-4, 5, (k, n)).astype(np.float32)
    A, B, C = cuda.Buffer.from_host(a), cuda.Buffer.from_host(b1), cuda.Buffer.alloc(m * n)

    def run():
        s0 = cuda.stats()["device_kernels"]
        cuda.matmul_3xtf32(A, B, C, m, n, k)
        return cuda.stats()["device_kernels"] - s0, C.to_host(m * n).reshape(m, n)

    want1 = (a.astype(np.float64) @ b1.astype(np.float64)).astype(np.float32)
    want2 = (a.astype(np.float64) @ b2.astype(np.float64)).astype(np.float32)
    cuda.set_operand_cache(True)
    k1, c1 = run()
    k2, c2 = run()
    assert (k1, k2) == (3, 2) and np.array_equal(c1, want1) and np.array_equal(c2, want1)
    B.upload(b2.ctypes.data, k * n)  # same buffer, new contents
    k3, c3 = run()
    k4, c4 = run()
    assert (k3, k4) == (3, 2) and np.array_equal(c3, want2) and np.array_equal(c4, want2)
    cuda.set_operand_cache(False)
    k5, c5 = run()
    k6, c6 = run()
    assert (k5, k6) == (3, 3) and np.array_equal(c6, want2)
    cuda.set_operand_cache(True)
    for x in (A, B, C):
        x.release()


def test_layout_is_not_symmetric(cuda):
    """catches transposed / swizzle-permuted operands that a random-sign test could hide"""
    m, n, k = 128, 256, 64
    a = (np.arange(m * k, dtype=np.float32).reshape(m, k) % 7) - 3
    b = (np.arange(k * n, dtype=np.float32).reshape(k, n) % 5) - 2
    assert np.array_equal(run_matmul(cuda, a, b), (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))
    e = np.zeros((m, k), np.float32)
    e[5, 9] = 1.0  # picks row 9 of b into row 5 of the result
    got = run_matmul(cuda, e, b)
    assert np.array_equal(got[5], b[9]) and not got[np.arange(m) != 5].any()


@pytest.mark.parametrize("m,n,k", [(256, 256, 512), (512, 768, 2048)])
def test_fp32_accuracy_3xtf32(cuda, m, n, k):
    a = finite_normal(m * k, 9).reshape(m, k)
    b = finite_normal(k * n, 10).reshape(k, n)
    got = run_matmul(cuda, a, b).astype(np.float64)
    truth = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    err = (np.abs(got - truth) / scale).max()
    assert err <= 1e-5, err  # north star; a single-pass TF32 product sits near 5e-4 here
    assert err <= 5e-6, err  # what 3xTF32 delivers with the tensor core's fp32 accumulation (measured 2.8e-6 at K = 2048)
    # the reference's own fp32 left fold, restated in C (oracle/oracle_cpu.c), is no closer to the truth
    from oracle import build as ob

    L = ob.load("strict")
    lf = np.empty((m, n), np.float32)
    L.oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, lf.ctypes.data, m, k, n)
    assert (np.abs(got - lf.astype(np.float64)) / scale).max() <= 1e-5


def accuracy(got, a, b):
    """max |got - truth| / (|A| . |B|), truth in fp64 — the scale BASELINE.md section 3 (C5, dataset N) states the 1e-5 bar on"""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    return float((np.abs(got.astype(np.float64) - a64 @ b64) / (np.abs(a64) @ np.abs(b64))).max())


@pytest.mark.parametrize("config", [1024, 512, 256, 128, 64])
@pytest.mark.parametrize("m,n,k", [(512, 768, 2048), (384, 520, 8192)])
def test_fp32_accuracy_on_normal_data_every_tile_configuration(cuda, monkeypatch, config, m, n, k):
    """dataset N (randomNormal: lo panels non-zero) on EVERY tile configuration incl. the CTA-pair kernel that carries the C5 number:
    a wrong lo tile, descriptor or half of B on CTA 1 shows up here as a TF32-sized (1e-4..1e-3) error"""
    monkeypatch.setenv("CC_GEMM_FORCE_CONFIG", str(config))
    a = finite_normal(m * k, 9).reshape(m, k)
    b = finite_normal(k * n, 10).reshape(k, n)
    got = run_matmul(cuda, a, b)
    err = accuracy(got, a, b)
    # north star 1e-5. The tensor core accumulates in fp32 with truncation, so the error grows linearly in K (2.8e-6 at K = 2048,
    # 5.1e-6 at K = 8192) where the reference's round-to-nearest left fold grows like sqrt(K)
    assert err <= (4e-6 if k <= 2048 else 7e-6), err
    from oracle import build as ob

    lf = np.empty((m, n), np.float32)
    ob.load("strict").oracle_matmul_left_fold(a.ctypes.data, b.ctypes.data, lf.ctypes.data, m, k, n)
    scale = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    assert (np.abs(got.astype(np.float64) - lf.astype(np.float64)) / scale).max() <= 1e-5  # vs the reference's fp32 left fold


@pytest.mark.parametrize("config", [1024, 512, 256, 128, 64])
@pytest.mark.parametrize("which", ["a_lo", "b_lo", "both"])
def test_each_cross_term_is_needed(cuda, monkeypatch, config, which):
    """lo-sensitivity: entries 1 + j * 2^-16 (j < 64) have hi = 1 and ALL their information in lo. With `a_lo` only A carries such entries
    (B is small integers, exactly TF32), so the result is right only if the A_lo . B_hi MMAs ran on the right tiles; `b_lo` is the mirror
    image for A_hi . B_lo. Dropping either cross term leaves an error of ~5e-4 relative, 100x the bar; lo . lo (dropped by design) is
    2^-22. Values depend on the position, so a swapped half of B on CTA 1 or a stale stage cannot cancel."""
    monkeypatch.setenv("CC_GEMM_FORCE_CONFIG", str(config))
    m, n, k = 512, 768, 512
    rng = np.random.default_rng(config + len(which))
    fine = lambda shape: (1.0 + rng.integers(0, 64, shape) * 2.0**-16).astype(np.float32)
    ints = lambda shape: rng.integers(1, 5, shape).astype(np.float32)
    a = fine((m, k)) if which in ("a_lo", "both") else ints((m, k))
    b = fine((k, n)) if which in ("b_lo", "both") else ints((k, n))
    got = run_matmul(cuda, a, b)
    truth = a.astype(np.float64) @ b.astype(np.float64)
    rel = float((np.abs(got.astype(np.float64) - truth) / truth).max())  # all entries positive: |A|.|B| = A.B
    assert rel <= 5e-6, rel  # fp32 accumulation over K = 512 (the tensor core's adds), far below the 3e-4 of a missing term
    # the same product with the lo information removed is far outside the bar, i.e. the assertion above has teeth
    hi_only = lambda x: (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    a_h = hi_only(a) if which in ("a_lo", "both") else a
    b_h = hi_only(b) if which in ("b_lo", "both") else b
    assert float((np.abs(a_h.astype(np.float64) @ b_h.astype(np.float64) - truth) / truth).max()) >= 2e-4


@pytest.mark.parametrize("config", [1024, 512, 128])
def test_k_chunks_keep_the_error_at_the_one_chunk_level(cuda, monkeypatch, config):
    """the tensor core's truncating fp32 accumulation makes one launch's error grow linearly in K (4.9e-6 at 8192): beyond 8192 the product
    runs in K chunks added in the epilogue with round-to-nearest, so K = 20000 (three chunks, the last one ragged) is no worse than one chunk;
    and integers stay exact through the C += chunk path"""
    monkeypatch.setenv("CC_GEMM_FORCE_CONFIG", str(config))
    m, n, k = 384, 264, 20000
    a = finite_normal(m * k, 9).reshape(m, k)
    b = finite_normal(k * n, 10).reshape(k, n)
    err = accuracy(run_matmul(cuda, a, b), a, b)
    assert err <= 6e-6, err
    rng = np.random.default_rng(config)
    ai = rng.integers(-4, 5, (m, k)).astype(np.float32)
    bi = rng.integers(-4, 5, (k, n)).astype(np.float32)
    assert np.array_equal(run_matmul(cuda, ai, bi), (ai.astype(np.float64) @ bi.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("m,n,k", [(1024, 1024, 1024), (512, 512, 4096), (700, 300, 1000), (1024, 518, 2052)])
@pytest.mark.parametrize("splits", [None, 1, 3, 8])
def test_split_k_on_small_products_is_exact_and_deterministic(cuda, monkeypatch, m, n, k, splits):
    """a product of few 256 x 256 tiles runs the tensor-memory-A kernel with K split over the CTA pairs, each (tile, split) unit writing its own
    partial result; the partials are added in split order by a second kernel: exact on integers, the same bits on every run, fp32-accurate on
    normal data; ragged M / N / K, N % 4 != 0 and uneven splits included (CC_GEMM_K_SPLITS forces a count, None = the launcher's own choice)"""
    monkeypatch.setenv("CC_GEMM_FORCE_CONFIG", "1024")
    if splits is not None:
        monkeypatch.setenv("CC_GEMM_K_SPLITS", str(splits))
    rng = np.random.default_rng(m + n + k)
    ai = rng.integers(-4, 5, (m, k)).astype(np.float32)
    bi = rng.integers(-4, 5, (k, n)).astype(np.float32)
    assert np.array_equal(run_matmul(cuda, ai, bi), (ai.astype(np.float64) @ bi.astype(np.float64)).astype(np.float32))
    a = finite_normal(m * k, 9).reshape(m, k)
    b = finite_normal(k * n, 10).reshape(k, n)
    first = run_matmul(cuda, a, b)
    assert accuracy(first, a, b) <= 4e-6
    for _ in range(3):
        assert np.array_equal(first.view(np.uint32), run_matmul(cuda, a, b).view(np.uint32))


@pytest.mark.parametrize("size", [1024, 768, 320])
def test_products_chained_through_their_outputs(cuda, size):
    """X <- X · P twenty times with P a permutation matrix and X, P re-written between products: every kernel of a product (split A, split B,
    contraction, sum of the K splits) is launched as a programmatic dependent of the one before it, so each may be resident while its input
    is still being written by its predecessor — the chain must still see exactly the previous product (1024: split K on the
    tensor-memory-A kernel; 768 / 320: operand panels + single-CTA tiles). Operand cache off: B is split again for every product."""
    rng = np.random.default_rng(size)
    x = rng.integers(-7, 8, (size, size)).astype(np.float32)
    perms = [rng.permutation(size) for _ in range(4)]
    cuda.set_operand_cache(False)
    try:
        X, Y = cuda.Buffer.from_host(x), cuda.Buffer.alloc(size * size)
        Ps = []
        for p in perms:
            m = np.zeros((size, size), np.float32)
            m[p, np.arange(size)] = 1.0  # (X · P)[:, j] = X[:, p[j]]
            Ps.append(cuda.Buffer.from_host(m))
        want = x
        for step in range(20):
            cuda.matmul_3xtf32(X, Ps[step % 4], Y, size, size, size)
            X, Y = Y, X
            want = want[:, perms[step % 4]]
        assert np.array_equal(X.to_host(size * size).reshape(size, size), want)
        for b in [X, Y] + Ps:
            b.release()
    finally:
        cuda.set_operand_cache(True)


def test_pattern_lowers_to_tcgen05(cuda):
    """matmul written the way benchmarks.scala:188-191 writes it runs on the tensor cores and never materialises i*j*k"""
    T = cuda.Tensor
    m, k, n = 256, 384, 512
    rng = np.random.default_rng(3)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    ta, tb = T(a), T(b)
    product = ta.broadcast([m, k, n]) * tb.reshape([1, k, n]).broadcast([m, k, n])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    kern = acc.compile()
    assert kern.info.kind == 2 and kern.info.n_args == 2 and kern.info.flops == 2 * m * n * k
    got = acc.flatArray().reshape(m, n)
    assert np.array_equal(got, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))


def matmul1(T, m1, m2):
    """benchmarks.scala:176-187 / README.md:312-329: join over the columns of m2 of left folds over the inner dimension"""
    cols1 = m1.split(1)
    outs = []
    for col2 in m2.split(1):
        terms = [l * r.broadcast(l.shape) for l, r in zip(cols1, col2.split(0))]
        acc = terms[0]
        for x in terms[1:]:
            acc = acc + x
        outs.append(acc)
    return T.join(outs)


@pytest.mark.parametrize("m,k,n,kind", [(1024, 128, 256, 2), (4096, 32, 32, 1), (300, 40, 36, 1), (64, 4, 8, 0)])
def test_matmul1_join_of_folds_is_rerolled(cuda, m, k, n, kind):
    T = cuda.Tensor
    rng = np.random.default_rng(m + k + n)
    a = rng.integers(-4, 5, (m, k)).astype(np.float32)
    b = rng.integers(-4, 5, (k, n)).astype(np.float32)
    e = matmul1(T, T(a), T(b))
    kern = e.compile()
    assert kern.info.kind == kind, kern.source[:300]
    assert kern.info.n_args == 2
    got = e.flatArray().reshape(m, n)
    assert np.array_equal(got, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))


def test_pattern_with_ragged_shape_runs_on_the_tensor_cores(cuda):
    T = cuda.Tensor
    m, k, n = 1000, 300, 700
    a = finite_normal(m * k, 9).reshape(m, k)
    b = finite_normal(k * n, 10).reshape(k, n)
    product = T(a).broadcast([m, k, n]) * T(b).reshape([1, k, n]).broadcast([m, k, n])
    parts = product.split(1)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    assert acc.compile().info.kind == 2
    got = acc.flatArray().reshape(m, n).astype(np.float64)
    scale = np.abs(a.astype(np.float64)) @ np.abs(b.astype(np.float64))
    assert (np.abs(got - a.astype(np.float64) @ b.astype(np.float64)) / scale).max() <= 5e-6


def test_full_size_8192_rows_sampled(cuda):
    """C5 at BASELINE size on dataset E (integers in {-4..4}: every order of summation is exact), checked on sampled rows
    and through the checksum identity sum(C) = colsum(A) . rowsum(B)"""
    n = 8192
    T = cuda.Tensor
    nine, four, one = T.fill(9.0, [n, n]), T.fill(4.0, [n, n]), T.fill(1.0, [n, n])

    def dataset_e(seed):
        r = T.random([n, n], seed=seed) * nine
        return ((r - r % one) - four).doCache()

    A, B = dataset_e(9), dataset_e(10)
    ab, bb, cb = A.doBuffer(), B.doBuffer(), cuda.Buffer.alloc(n * n)
    cuda.matmul_3xtf32(ab, bb, cb, n, n, n)
    a = ab.to_host().reshape(n, n)
    b = bb.to_host().reshape(n, n)
    c = cb.to_host().reshape(n, n)
    rows = np.r_[0:8, 127:130, 4095:4098, 8184:8192]
    want = a[rows].astype(np.float64) @ b.astype(np.float64)
    assert np.array_equal(c[rows].astype(np.float64), want)
    total = float(c.astype(np.float64).sum())
    assert total == float(a.astype(np.float64).sum(axis=0) @ b.astype(np.float64).sum(axis=1))
    for x in (ab, bb, cb):
        x.release()


@pytest.mark.parametrize("b,m,k,n", [(3, 128, 96, 256), (8, 200, 130, 72), (2, 512, 512, 512)])
def test_batched_matmul_lowers_to_one_pipeline_launch_per_batch(cuda, monkeypatch, b, m, k, n):
    """C[b, i, k] = sum_t A[b, i, t] * B[b, t, k] (matmul2 with a leading batch dim): batch dims go into the rows of both operand
    panels and the tcgen05 pipeline runs once per batch on its block of rows.  Exact on small integers.  The panel gathers and the
    per-batch blocking are checked on the CPU tier (tests/test_kernel_emulation.py); this is the device half."""
    monkeypatch.setenv("CC_TUNE_CONTRACTION_MIN_MACS", "1")  # (the default threshold keeps the first two shapes on the generic reduction)
    cuda.kernel_cache_clear()
    try:
        rng = np.random.default_rng(b * 1000 + m)
        A = rng.integers(-4, 5, (b, m, k)).astype(np.float32)
        B = rng.integers(-4, 5, (b, k, n)).astype(np.float32)
        T = cuda.Tensor
        a4 = T(A).broadcast([b, m, k, n])
        b4 = T(B).reshape([b, 1, k, n]).broadcast([b, m, k, n])
        parts = (a4 * b4).split(2)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        kern = acc.compile()
        assert kern.info.kind == 2 and f"(batch of {b})" in kern.source
        got = acc.flatArray().reshape(b, m, n)
        want = np.einsum("bmk,bkn->bmn", A.astype(np.int64), B.astype(np.int64)).astype(np.float32)
        assert np.array_equal(got, want)
    finally:
        cuda.kernel_cache_clear()
