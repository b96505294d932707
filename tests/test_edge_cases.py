"""GPU parity, part 3: the edges of the path — empty and rank-0 tensors, unit dimensions, ranks above 3 (the reference folds
the leading dimensions into global id 0, OpenCLKernelBuilder.scala:177-211), views that leave the source entirely, fractional
offsets (the `(int)` truncation quirk, OpenCLKernelBuilder.scala:386), non-zero paddings, and index spaces beyond 2^31 elements
(the reference's `shape.product` is an Int, Tensors.scala:1345; the CUDA path switches its index type).  Small cases are
compared bit for bit with the CPU oracle; the large ones through properties."""
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


def same(cuda, build):
    got = build(cuda.Tensor)
    want = build(ref.Tensor)
    assert tuple(got.shape) == tuple(want.shape)
    g, w = got.flatArray(), want.flat_array()
    assert np.array_equal(bits(g), bits(w)), (g, w)
    assert got.toString() == want.to_string()
    return g


def test_rank0(cuda):
    same(cuda, lambda T: T.scalar(3.0))
    same(cuda, lambda T: T.scalar(3.0) * T.scalar(0.5) + T.scalar(1.0))
    same(cuda, lambda T: T.scalar(3.0).transpose())
    same(cuda, lambda T: T.scalar(3.0).broadcast([2, 3]))
    same(cuda, lambda T: T.scalar(-2.5).broadcast([4, 1, 3]) * T.random([4, 1, 3], seed=1))
    same(cuda, lambda T: T.scalar(7.0).sum())
    same(cuda, lambda T: T.join([T.scalar(1.0), T.scalar(2.0), T.scalar(3.0)]))
    same(cuda, lambda T: T.random([5], seed=2).split(0)[3])
    same(cuda, lambda T: T.random([5], seed=2).split(0)[3].broadcast([2, 2]) + T.fill(1.0, [2, 2]))


def test_unit_dimensions(cuda):
    same(cuda, lambda T: T.random([1], seed=1))
    same(cuda, lambda T: T.random([1, 1, 1], seed=1).permute([2, 0, 1]))
    same(cuda, lambda T: T.random([3, 1, 4], seed=1).split(1)[0])
    same(cuda, lambda T: T.join(T.random([3, 1, 4], seed=1).split(1)))
    same(cuda, lambda T: T.join([T.random([2, 3], seed=1)]))
    same(cuda, lambda T: T.join([T.random([2, 3], seed=1)], 0))
    same(cuda, lambda T: T.random([1, 7], seed=1).broadcast([1, 7, 1]))
    same(cuda, lambda T: T.random([7, 1], seed=1).transpose() + T.random([1, 7], seed=2))
    same(cuda, lambda T: T.random([1, 5], seed=1).sum())
    x = lambda T: T.random([1, 6], seed=3)  # noqa: E731
    same(cuda, lambda T: x(T).split(0)[0] * T.fill(2.0, [6]))


def test_ranks_above_three(cuda):
    """leading dims folded into global id 0 by the reference (K:177-211); here: one linearised index space"""
    same(cuda, lambda T: T.random([2, 3, 4, 5], seed=1) * T.random([2, 3, 4, 5], seed=2))
    same(cuda, lambda T: T.random([2, 3, 4, 5], seed=1).permute([3, 1, 0, 2]))
    same(cuda, lambda T: T.random([2, 3, 4, 5, 6], seed=1).permute([4, 0, 3, 1, 2]).translate([1, 0, -1, 2, 0]))
    same(cuda, lambda T: T.random([3, 4], seed=1).broadcast([3, 4, 2, 5, 2]))
    same(cuda, lambda T: T.random([2, 3, 4, 5], seed=1).translate([0, 1, -2, 3]) + T.random([2, 3, 4, 5], seed=2).permute([0, 1, 2, 3]))
    same(cuda, lambda T: T.join(T.random([2, 3, 4, 5], seed=1).split(2), 1))
    same(cuda, lambda T: T.random([2, 2, 2, 2, 2, 2], seed=9).permute([5, 4, 3, 2, 1, 0]))
    # 64 x 64 tiles engage on the big-enough version of the same thing
    same(cuda, lambda T: T.random([2, 70, 3, 66], seed=1).permute([0, 3, 2, 1]))
    g = same(cuda, lambda T: T.random([3, 5, 7, 9], seed=4).sum())
    assert g.shape == (1,)


def test_join_at_every_dimension(cuda):
    """Tensor.join(tensors, dimension) (T:560-575): the reference joins last and gathers a permuted view (two kernels); here the
    join kernel stores with the element index at `dimension` — re-rolled joins, tuple-store joins, joins of reductions"""
    for d in (0, 1, 2, 3):
        same(cuda, lambda T: T.join(T.random([3, 4, 5, 6], seed=1).split(d), d))          # round trip = identity
        same(cuda, lambda T: T.join(T.random([3, 4, 5, 6], seed=1).split((d + 1) % 4), d))  # moves a dimension
    for d in (0, 1, 2):
        # unrelated elements: per-index stores
        same(cuda, lambda T: T.join([T.abs(-T.random([4, 6], seed=1)), T.random([4, 6], seed=2) * T.random([4, 6], seed=3), T.fill(2.0, [4, 6])], d))
        # elements that are per-axis sums (matmul1-style join of folds)
        def folds(T):
            r = T.random([9, 5, 8], seed=4) * T.fill(9.0, [9, 5, 8])
            x = r - r % T.fill(1.0, [9, 5, 8])  # integers 0..8: sums are exact in any order
            cols = []
            for part in x.split(1):
                rows = part.split(0)
                acc = rows[0]
                for r in rows[1:]:
                    acc = acc + r
                cols.append(acc)
            return T.join(cols, d) if d < 2 else T.join(cols)
        same(cuda, folds)
    same(cuda, lambda T: T.join([T.scalar(1.0), T.scalar(2.0)], 0))
    same(cuda, lambda T: T.join([T.random([7], seed=1), T.random([7], seed=2)], 0))
    same(cuda, lambda T: T.join([T.random([7], seed=1), T.random([7], seed=2)], 0).translate([0, 1]) * T.fill(2.0, [2, 7]))
    # big enough for the vector / tiled-transpose templates
    g = same(cuda, lambda T: T.join(T.random([64, 96, 128], seed=5).split(0), 2))
    assert g.shape == (64 * 96 * 128,)
    k = cuda.Tensor.join(cuda.Tensor.random([64, 96, 128], seed=5).split(1), 1).compile()
    assert k.info.kind == 0 and "flat=1" in k.source  # split(d) joined back at d is the identity copy, one kernel


def test_dense_windows_through_the_stencil_tile(cuda):
    """box sums / max pooling written as translated views of one source: staged through shared memory (tile + halo, bounds and
    padding applied while staging). Integer-valued data and max keep every result exact -> bit for bit vs the oracle."""
    def data(shape, seed):
        return (np.floor(ref.random_buffer(int(np.prod(shape)), seed) * np.float32(9.0)) - np.float32(4.0)).astype(np.float32).reshape(shape)

    def fold(terms, f):
        acc = terms[0]
        for t in terms[1:]:
            acc = f(acc, t)
        return acc

    cases = []
    for shape, pad in (([40, 256], 0.0), ([17, 128], -2.0), ([2, 3, 37, 132], 5.0), ([64, 384], -7.0)):
        rank = len(shape)
        lead = [0] * (rank - 2)
        x_np, y_np = data(shape, 3), data(shape, 4)

        def window(T, offsets, f, lead=lead, extra=False, x_np=x_np, y_np=y_np, pad=pad, rank=rank):
            x = T(x_np, padding=pad)
            e = fold([x.translate(lead + [dy, dx]) for dy, dx in offsets], f)
            return e * T(y_np) + T(y_np).translate([0] * (rank - 2) + [1, 0]) if extra else e

        w3 = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        w5 = [(dy, dx) for dy in range(-2, 3) for dx in range(-2, 3)]
        skew = [(dy, dx) for dy in (0, 1, 2) for dx in (-5, -1, 0, 3)]
        row6 = [(0, dx) for dx in (-3, -2, -1, 1, 2, 3)]
        cases += [
            (lambda T, o=w3, w=window: w(T, o, lambda a, b: a + b), "3x3 sum"),
            (lambda T, o=w3, w=window: w(T, o, T.max), "3x3 max"),
            (lambda T, o=w5, w=window: w(T, o, lambda a, b: a + b, extra=True), "5x5 sum with other operands"),
            (lambda T, o=skew, w=window: w(T, o, T.min), "skewed window"),
            (lambda T, o=row6, w=window: w(T, o, lambda a, b: a + b), "1-D window of 6"),
        ]
        if rank > 2:
            cases.append((lambda T, o=w3, w=window, rank=rank: w(T, o, T.max, lead=[1] + [0] * (rank - 3)), "window shifted along a leading dimension"))
    for build, name in cases:
        k = build(cuda.Tensor).compile()
        assert "stencil tile" in k.source, name
        same(cuda, build)
    # too small / too few views: the ordinary templates
    x = cuda.Tensor.random([16, 64], seed=1)
    assert "stencil tile" not in fold([x.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], cuda.Tensor.max).compile().source
    y = cuda.Tensor.random([64, 256], seed=1)
    assert "stencil tile" not in (y + y.translate([0, 1])).compile().source


def test_views_that_leave_the_source(cuda):
    for pad in (0.0, -1.5, float("inf")):
        same(cuda, lambda T: T.random([4, 6], seed=1, padding=pad).translate([4, 0]))       # entirely padding
        same(cuda, lambda T: T.random([4, 6], seed=1, padding=pad).translate([-7, 9]))
        same(cuda, lambda T: T.random([4, 6], seed=1, padding=pad).translate([3, -5]))      # one element survives
        same(cuda, lambda T: T.random([4, 6], seed=1, padding=pad).translate([1, 1], [9, 11]))  # grown output shape
        same(cuda, lambda T: T.random([4, 8], seed=1, padding=pad).translate([0, 2]) * T.random([4, 8], seed=2, padding=pad).translate([0, -2]))
    # the padding of the RESULT is the receiver's padding (T:970-1003): a second view pads with it again
    same(cuda, lambda T: T.random([4, 8], seed=1, padding=2.0).translate([1, 1]).translate([-2, -2]))
    # ... while an inline expression's padding is the same function of the operands' paddings (T:857-863)
    same(cuda, lambda T: (T.random([4, 8], seed=1, padding=2.0) * T.fill(3.0, [4, 8], padding=3.0)).translate([1, 3]))


def test_fractional_offsets_truncate_toward_zero(cuda):
    """K:386 `(int)` conversion: an index in (-1, 0) lands on element 0, not on the padding"""
    same(cuda, lambda T: T.random([2, 3], seed=1).translate([0.5, -0.25]))
    same(cuda, lambda T: T.random([5], seed=1).translate([-0.999]))
    same(cuda, lambda T: T.random([5], seed=1).translate([1.5]))
    same(cuda, lambda T: T.random([6, 4], seed=1).translate([0.5, 0]).permute([1, 0]))
    same(cuda, lambda T: T.random([3, 3], seed=1).scale([7, 5]))
    same(cuda, lambda T: T.random([8, 8], seed=1).scale([3, 3]) + T.random([3, 3], seed=2))


def test_empty_tensors(cuda):
    """a zero-sized global work size is an error in OpenCL 1.x, so the reference has no behaviour to match; here an empty tensor
    evaluates to an empty array without a launch, and folds to the monoid's zero"""
    T = cuda.Tensor
    assert T.fill(1.0, [0]).flatArray().shape == (0,)
    assert T.random([3, 0, 2], seed=1).flatArray().shape == (0,)
    assert (T.random([3, 0, 2], seed=1) * T.fill(2.0, [3, 0, 2])).flatArray().shape == (0,)
    assert T.tanh(T.random([0, 4], seed=1)).permute([1, 0]).shape == (4, 0)
    assert T.tanh(T.random([0, 4], seed=1)).permute([1, 0]).flatArray().shape == (0,)
    assert T.fill(1.0, [0, 4]).sum().flatArray().tolist() == [0.0]
    assert T.random([0], seed=1).sum().flatArray().tolist() == [0.0]
    assert T.random([0], seed=1).reduce("*").flatArray().tolist() == [1.0]
    assert T.random([2, 0], seed=1).toString() == ref.Tensor.fill(0.0, [2, 0]).to_string()
    with T.fill(1.0, [0]).flatBuffer() as a:
        assert a.shape == (0,)


def test_index_space_beyond_2_31(cuda):
    """2^31 + 2^20 elements (8.6 GB per tensor): flat elementwise, a broadcast view, a transpose, the whole-tensor fold"""
    T = cuda.Tensor
    rows, cols = 32768 + 16, 65536
    n = rows * cols
    assert n > 2**31
    x = T.random([rows, cols], seed=3).doCache()
    # the generator itself is 32-bit by construction (hash of a uint index, T:432-443): check the far end against the oracle
    tail = x.doBuffer()
    far = tail.to_host(4096, offset=n - 4096)
    i = np.arange(n - 4096, n, dtype=np.uint64).astype(np.uint32)
    want_far = (ref.wang_hash(i ^ np.uint32(3)).astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)
    assert np.array_equal(bits(far), bits(want_far))
    tail.release()
    # flat elementwise with a 64-bit index
    e = x * T.fill(2.0, [rows, cols]) + T.fill(1.0, [rows, cols])
    k = e.compile()
    assert "idx=long long" in k.source
    b = e.doBuffer()
    got = b.to_host(4096, offset=n - 4096)
    assert np.array_equal(bits(got), bits(want_far * np.float32(2.0) + np.float32(1.0)))
    got0 = b.to_host(4096, offset=0)
    i0 = np.arange(4096, dtype=np.uint32)
    want0 = (ref.wang_hash(i0 ^ np.uint32(3)).astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)
    assert np.array_equal(bits(got0), bits(want0 * np.float32(2.0) + np.float32(1.0)))
    b.release()
    # whole-tensor fold: uniform[0,1) mean 1/2
    s = float(x.sum().flatArray()[0])
    assert abs(s / n - 0.5) < 1e-4
    # a view whose source offsets exceed 2^31: last row broadcast down a short tensor, and the transpose of the far corner
    last_row = x.translate([-(rows - 1), 0], [1, cols])
    assert np.array_equal(bits(last_row.flatArray()), bits(_rows_of(3, rows - 1, 1, cols)))
    corner = x.translate([-(rows - 64), -(cols - 64)], [64, 64]).permute([1, 0])
    want_corner = _block(3, rows - 64, cols - 64, 64, 64, cols).T
    assert np.array_equal(bits(corner.flatArray()), bits(np.ascontiguousarray(want_corner).ravel()))
    # full-size transposed view with 64-bit offsets, sampled
    t = x.permute([1, 0])
    tb = t.doBuffer()
    got_t = tb.to_host(rows, offset=(cols - 1) * rows)  # last row of the transpose = last column of x
    want_t = _block(3, 0, cols - 1, rows, 1, cols).ravel()
    assert np.array_equal(bits(got_t), bits(want_t))
    tb.release()


def _block(seed, r0, c0, nr, nc, cols):
    r = np.arange(r0, r0 + nr, dtype=np.uint64)[:, None]
    c = np.arange(c0, c0 + nc, dtype=np.uint64)[None, :]
    i = (r * np.uint64(cols) + c).astype(np.uint32)
    return (ref.wang_hash(i ^ np.uint32(seed)).astype(np.float32) / np.float32(4294967296.0)).astype(np.float32)


def _rows_of(seed, r0, nr, cols):
    return _block(seed, r0, 0, nr, cols, cols).ravel()
