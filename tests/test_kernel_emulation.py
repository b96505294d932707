"""CPU-only: generated kernels run on the host emulator (tests/kernel_emulator, test infrastructure only) against the oracle.
Exactly the CUDA text NVRTC compiles for sm_100a is compiled with g++ instead and executed with one host thread per CUDA thread,
so the code generator's indexing, bounds tests, padding, lane picks, tiles, shuffles and fold orders are exercised where no GPU
exists.  Integer-valued data keeps every case bit-exact whatever the evaluation order.  (The GPU tier repeats all of this on the
device; this tier makes no parity or performance claim.)"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "kernel_emulator"))
from run import arg_ordinals, emulate  # noqa: E402

from compute.scala_b200 import cuda  # noqa: E402
from oracle import reference as ref  # noqa: E402

T, R = cuda.Tensor, ref.Tensor


def ints(B, shape, seed, padding=0.0):
    """integers in {-4..4} from the bit-reproducible hash RNG, materialised as a leaf on either backend (lazy on the cuda side)"""
    r = B.random(shape, seed=seed, padding=padding) * B.fill(9.0, shape)
    e = r - r % B.fill(1.0, shape) - B.fill(4.0, shape)
    return e


def oracle_params(rt, root_is_join=True):
    if root_is_join and isinstance(rt, ref._Transformed) and isinstance(rt.checkpoint, ref._Join):  # join(…, dimension) = permuted view of the last-dim join (T:560-575)
        rt = rt.checkpoint
    if root_is_join and isinstance(rt, ref._Join):
        seen, out = set(), []
        for t in rt._tensors:
            for p in ref.parameter_descendants(t.closure()):
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
        return out
    if isinstance(rt, ref._Sum):
        return ref.parameter_descendants(rt._base.closure())
    return ref.parameter_descendants(rt.closure())


def check(build, expect_in_source=None, kind=None):
    """build(B, leaf) -> expression; leaf(shape, seed, padding) makes an integer-valued NON-INLINE leaf on backend B"""
    def leaf_r(shape, seed, padding=0.0):
        data = ints(R, shape, seed).flat_array().reshape(shape)
        return R(data, padding=padding)

    def leaf_t(shape, seed, padding=0.0):
        return T.random(shape, seed=seed, padding=padding)  # stands in for the leaf: only its position in the argument list matters

    want_t = build(R, leaf_r)
    want = want_t.flat_array()
    expr = build(T, leaf_t)
    params = oracle_params(want_t)
    k0 = expr.compile()
    ords = arg_ordinals(cuda, k0)
    if len(params) == 1 and ords and min(ords) >= 1:
        # views of ONE unevaluated inline tensor (matmul2: `P.split(1).reduce(_ + _)`): the cuda side composes P's closure into the kernel,
        # so its arguments are P's own leaves, numbered after the main tree's parameter (ordinal 0 = P, which it never materialises)
        inner = ref.parameter_descendants(params[0].id.closure())
        leaves = [inner[o - 1].id.buffer() for o in ords]
    else:
        assert all(0 <= o < len(params) for o in ords), (ords, len(params))
        leaves = [params[o].id.buffer() for o in ords]
    got, k = emulate(cuda, expr, leaves)
    if expect_in_source:
        assert expect_in_source in k.source, expect_in_source
    if kind is not None:
        assert k.info.kind == kind
    assert tuple(expr.shape) == tuple(want_t.shape)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or np.array_equal(got, want), (got[:16], want[:16])


def chain(parts, f=lambda a, b: a + b):
    acc = parts[0]
    for p in parts[1:]:
        acc = f(acc, p)
    return acc


def test_elementwise_vector_and_scalar_lanes():
    check(lambda B, leaf: B.abs(leaf([12, 16], 1) * leaf([12, 16], 2) - leaf([12, 16], 3)), "flat=1", 0)
    check(lambda B, leaf: B.max(leaf([5, 7], 1), -leaf([5, 7], 2)) + leaf([5, 7], 3), None, 0)            # 35 elements: vector body + scalar tail
    check(lambda B, leaf: leaf([6, 9], 1, 2.0).translate([1, -2]) * leaf([6, 9], 2), "V=1", 0)             # odd fastest dimension under a view
    check(lambda B, leaf: leaf([3], 1).broadcast([3, 8]) + leaf([3, 8], 2), None, 0)                       # trailing broadcast: one scalar feeds 4 lanes
    check(lambda B, leaf: leaf([2, 8], 1).reshape([1, 2, 8]).broadcast([3, 2, 8]) - leaf([3, 2, 8], 2), None, 0)


def test_translations_paddings_and_shifted_vector_loads():
    for off in ([0, 1], [0, -1], [1, 3], [-2, 4], [0, 5], [3, 0]):
        check(lambda B, leaf, off=off: leaf([6, 16], 1, -2.0).translate(off), None, 0)
    # several shifted views of one narrow source: aligned vector pairs with static lane picks and clamped addresses
    def window(B, leaf):
        x = leaf([9, 32], 4, 3.0)
        return chain([x.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)], B.max)
    check(window, "float A", 0)
    check(lambda B, leaf: leaf([4, 4, 8], 1, 1.0).translate([1, 0, -1]).translate([0, -1, 2]), None, 0)


def test_ranks_zero_to_six_and_non_integer_coefficients():
    # rank 0 (a scalar kernel: `*out = term`, K:149-161) and unit dimensions
    check(lambda B, leaf: B.abs(leaf([], 1)) + B.scalar(2.0), None, 0)
    check(lambda B, leaf: leaf([1, 5, 1], 1) * leaf([1, 5, 1], 2), None, 0)
    # ranks 5 and 6: more than three dimensions fold into the flat index space (K:177-211)
    check(lambda B, leaf: leaf([2, 3, 2, 3, 4], 1).permute([4, 0, 3, 1, 2]).translate([1, 0, -1, 0, 1]), None, None)
    check(lambda B, leaf: leaf([2, 2, 3, 2, 2, 4], 1, -1.0).translate([0, 1, 0, -1, 0, 2]) + leaf([2, 2, 3, 2, 2, 4], 2), None, 0)
    # fractional offsets: DecimalFormat rounding of the coefficient, double arithmetic, (int) truncation toward zero (K:14-32, 372-386)
    for off in ([0.5, -1.5], [-0.25, 2.75], [1.0006, -0.9994]):
        check(lambda B, leaf, off=off: leaf([6, 8], 1, 9.0).translate(off), None, 0)
    # scale: the view matrix has non-integer diagonal coefficients (T:950-965)
    check(lambda B, leaf: leaf([4, 6], 1, 7.0).scale([6, 9]), None, 0)
    check(lambda B, leaf: leaf([5, 8], 1).scale([3, 4]) * leaf([3, 4], 2), None, 0)


def test_tiled_transposes():
    check(lambda B, leaf: leaf([40, 36], 1).transpose(), "tiled transpose", 3)
    check(lambda B, leaf: leaf([3, 34, 33], 2).permute([0, 2, 1]) + leaf([3, 33, 34], 3), "tiled transpose", 3)
    check(lambda B, leaf: leaf([33, 2, 40], 2, 5.0).permute([2, 1, 0]).translate([1, 0, -1]), "tiled transpose", 3)
    check(lambda B, leaf: B.join(leaf([34, 3, 36], 5).split(1)), None, None)  # join of a split: re-rolled into an output dimension


def test_stencil_tile():
    def window(B, leaf, shape=(20, 132), pad=-2.0, lead=()):
        x = leaf(list(shape), 4, pad)
        offs = [(dy, dx) for dy in (-1, 0, 2) for dx in (-5, -1, 0, 3)]
        e = chain([x.translate(list(lead) + [dy, dx]) for dy, dx in offs])
        return e * leaf(list(shape), 5) - leaf(list(shape), 6).translate([0] * (len(shape) - 1) + [1])
    check(window, "stencil tile", 0)
    check(lambda B, leaf: window(B, leaf, (2, 9, 128), 0.0, (1,)), "stencil tile", 0)


def test_axis_reductions_every_owner_and_monoid():
    check(lambda B, leaf: chain(leaf([12, 128], 1).split(0)), "column owner", 1)
    check(lambda B, leaf: chain(leaf([12, 40], 1).split(0)), "row owner", 1)                          # < 64 outputs: a warp per output
    check(lambda B, leaf: chain(leaf([300, 64], 1).split(0)), "column owner", 1)                     # T split over blockIdx.y + second stage
    check(lambda B, leaf: chain(leaf([9, 64], 1).split(1)), "row owner", 1)
    check(lambda B, leaf: chain(leaf([6, 2048], 1).split(1)), "threads/output=256", 1)
    check(lambda B, leaf: chain(leaf([12, 128], 1).split(0), B.max), "fold=Max", 1)
    check(lambda B, leaf: chain(leaf([9, 64], 1).split(1), B.min), "fold=Min", 1)
    check(lambda B, leaf: B.fill(2.0, [128]) * chain(leaf([12, 128], 1).split(0)) - leaf([128], 2), "epilogue=1", 1)
    # a nested (3 x 3 x channel) weighted window: the convolution idiom, re-rolled over three digits
    def conv(B, leaf):
        x, w = leaf([2, 6, 8, 3], 1), leaf([3, 3, 3], 2)
        terms = [x.split(3)[c].translate([0, dy - 1, dx - 1]) * w.split(0)[dy].split(0)[dx].split(0)[c].broadcast([2, 6, 8]) for dy in range(3) for dx in range(3) for c in range(3)]
        return chain(terms)
    check(conv, "T=3x3x3", 1)


def test_general_contraction_panels_and_epilogue(monkeypatch):
    """the implicit-GEMM lowering (convolution; matmul over views) with the size threshold lowered so that it is reached at emulator
    sizes: generated panel gathers (bounds tests, zero padding, K padded to 32, exact TF32 hi parts) and the in-place epilogue run on the
    host, a float64 product of the panels stands in for the tcgen05 pipeline"""
    monkeypatch.setenv("CC_TUNE_CONTRACTION_MIN_MACS", "1")
    cuda.kernel_cache_clear()
    try:
        def conv(B, leaf, depth=4, filters=32):
            x, w, bias = leaf([2, 5, 6, depth], 1), leaf([3, 3, depth, filters], 2), leaf([filters], 3)
            xs = x.split(3)
            ws = [[[wc.split(0) for wc in wx.split(0)] for wx in wy.split(0)] for wy in w.split(0)]  # ws[dy][dx][c][f]: scalars
            bs = bias.split(0)
            outs = []
            for f in range(filters):
                terms = [xs[c].translate([0, dy - 1, dx - 1]) * ws[dy][dx][c][f].broadcast([2, 5, 6]) for dy in range(3) for dx in range(3) for c in range(depth)]
                outs.append(chain(terms) + bs[f].broadcast([2, 5, 6]))
            return B.join(outs)
        check(conv, "over gathered operand panels", 2)

        def matmul_over_views(B, leaf):  # A stored transposed, B shifted by one row (zero padding enters the panel)
            at, b = leaf([40, 36], 4), leaf([40, 64], 5)
            a3 = at.transpose().broadcast([36, 40, 64])                     # A[i, t] broadcast over k
            b3 = b.translate([1, 0]).reshape([1, 40, 64]).broadcast([36, 40, 64])
            return chain((a3 * b3).split(1))
        check(matmul_over_views, "over gathered operand panels", 2)

        # batched matmul C[b, i, k] = sum_t A[b, i, t] * B[b, t, k]: on the tensor-core pipeline by default since round 2 (profiles/r02_knob_ab.json)
        def batched(B, leaf):
            a, b = leaf([3, 36, 40], 6), leaf([3, 40, 64], 7)
            a4 = a.broadcast([3, 36, 40, 64])
            b4 = b.reshape([3, 1, 40, 64]).broadcast([3, 36, 40, 64])
            return chain((a4 * b4).split(2))
        check(batched, "(batch of 3)", 2)
        monkeypatch.setenv("CC_BATCHED_CONTRACTION", "0")
        cuda.kernel_cache_clear()
        check(batched, "column owner", 1)  # switched off: the generic re-rolled reduction
        monkeypatch.delenv("CC_BATCHED_CONTRACTION")
    finally:
        monkeypatch.delenv("CC_TUNE_CONTRACTION_MIN_MACS")
        cuda.kernel_cache_clear()


def test_fused_second_stage_of_a_split_axis_reduction(monkeypatch):
    """default since round 2 (profiles/r02_knob_ab.json): the last CTA to finish a block of columns folds that block's partials inside
    reduce_cols -- one launch instead of two; the block counters reset themselves (checked by the emulator after every run)"""
    cuda.kernel_cache_clear()
    try:
        for _ in range(2):  # twice: the second run starts from the counters the first one left
            check(lambda B, leaf: chain(leaf([300, 64], 1).split(0)), "second stage fused into reduce_cols", 1)
        check(lambda B, leaf: chain(leaf([200, 260], 1).split(0)), "second stage fused into reduce_cols", 1)       # ragged column block
        check(lambda B, leaf: chain(leaf([300, 66], 1).split(0), B.max), "second stage fused into reduce_cols", 1)  # V = 1 lanes, max
        check(lambda B, leaf: B.abs(chain(leaf([300, 64], 1).split(0))) - leaf([64], 2), "second stage fused into reduce_cols", 1)  # epilogue in stage 2
        k = chain(T.random([300, 64], seed=1).split(0)).compile()
        assert k.info.n_launches == 1
        monkeypatch.setenv("CC_FUSE_COL_STAGE", "0")
        cuda.kernel_cache_clear()
        check(lambda B, leaf: chain(leaf([300, 64], 1).split(0)), "column owner", 1)  # switched off: reduce_cols + reduce_partials
        k = chain(T.random([300, 64], seed=1).split(0)).compile()
        assert k.info.n_launches == 2 and "fused into reduce_cols" not in k.source
    finally:
        monkeypatch.delenv("CC_FUSE_COL_STAGE", raising=False)
        cuda.kernel_cache_clear()


def convolute(B, inp, weight, bias):
    """benchmarks.scala:463-556"""
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
        outs.append(bias_seq[f].broadcast([batch, height, width]) + chain(summands))
    return B.join(outs)


def test_tile_owner_small_trailing_output_dimension(monkeypatch):
    """the reference's own benchmark shapes (benchmarks.scala:612-630: 3 x 3 and 1 x 1 kernels, depth 8 / 3; :196-198 skinny products): a thread
    owns every output along the small trailing dimension, lane-invariant loads are shared by all of them and fetched as one vector along the
    innermost reduction digit where that is contiguous"""
    monkeypatch.setenv("CC_REDUCE_TILE_OWNER", "2")  # (the shapes the emulator can afford are below the size from which it is the default)
    cuda.kernel_cache_clear()
    check(lambda B, leaf: convolute(B, leaf([3, 9, 10, 8], 1), leaf([3, 3, 8, 8], 2), leaf([8], 3)), "tile owner): out dims=[3,9,10,8] T=3x3x8 F=8 (2 vectors per thread) P=2 kvec=4", 1)
    check(lambda B, leaf: convolute(B, leaf([3, 9, 10, 8], 1), leaf([1, 1, 8, 8], 2), leaf([8], 3)), "tile owner", 1)
    check(lambda B, leaf: convolute(B, leaf([2, 5, 12, 8], 1), leaf([3, 3, 8, 8], 2), leaf([8], 3)), "P=4 kvec=4", 1)  # four positions per thread, window across the tile edge
    check(lambda B, leaf: convolute(B, leaf([2, 5, 12, 4], 1), leaf([3, 5, 4, 4], 2), leaf([4], 3)), "P=4 kvec=4", 1)  # 3 x 5 window
    # depth 3 (no vector along the channels), 12 filters = 3 vectors per thread
    check(lambda B, leaf: convolute(B, leaf([2, 7, 9, 3], 1), leaf([3, 3, 3, 12], 2), leaf([12], 3)), "F=12 (3 vectors per thread) P=1 kvec=1", 1)
    # depth 3, 3 filters: scalar lanes, the column owner as before
    check(lambda B, leaf: convolute(B, leaf([2, 7, 9, 3], 1), leaf([3, 3, 3, 3], 2), leaf([3], 3)), "column owner", 1)

    def skinny(B, leaf, m=300, k=16, n=8, fold=None):  # benchmarks.scala:188-191 on a tall A
        a, b = leaf([m, k], 1), leaf([k, n], 2)
        prod = a.broadcast([m, k, n]) * b.reshape([1, k, n]).broadcast([m, k, n])
        return chain(prod.split(1), fold) if fold else chain(prod.split(1))
    check(skinny, "tile owner", 1)
    check(lambda B, leaf: skinny(B, leaf, k=10, n=32), "F=32 (8 vectors per thread) P=1 kvec=1", 1)
    check(lambda B, leaf: skinny(B, leaf, fold=B.max), "fold=Max", 1)
    monkeypatch.delenv("CC_REDUCE_TILE_OWNER")
    cuda.kernel_cache_clear()


def test_small_n_contraction_on_emulated_warp_mmas():
    """the small-N contraction kernel (warp-level m16n8k8 MMAs, B as TF32 hi / lo fragments in registers or shared memory, A gathered through
    its affine map) with the MMA emulated for the 32 host threads of a warp: the implicit im2col (window offsets, padding at the image
    border), K padded to a multiple of 8, rows not a multiple of 16, the bias epilogue on the accumulator fragments, both B placements and
    both A load widths — exact on exactly representable data"""
    check(lambda B, leaf: convolute(B, leaf([4, 32, 32, 8], 1), leaf([3, 3, 8, 8], 2), leaf([8], 3)), "small-N contraction 4096x8x72", 1)         # B in registers, scalar A loads
    check(lambda B, leaf: convolute(B, leaf([1, 64, 65, 4], 1), leaf([3, 3, 4, 16], 2), leaf([16], 3)), "small-N contraction 4160x16x36", 1)      # K = 36 padded to 40, two n tiles: pair loads
    check(lambda B, leaf: convolute(B, leaf([1, 65, 64, 16], 1), leaf([3, 3, 16, 16], 2), leaf([16], 3)), "B in shared memory", 1)               # 36 fragments

    def skinny(B, leaf, m=4101, k=28, n=12):  # rows % 16 != 0, n % 8 != 0, k % 8 != 0; A stored transposed and shifted along k with a padding
        at, b = leaf([k, m], 1, padding=2.5), leaf([k, n], 2)
        a = at.permute([1, 0]).translate([0, 1])  # [m, k]: column 0 is the padding
        prod = a.broadcast([m, k, n]) * b.reshape([1, k, n]).broadcast([m, k, n])
        return chain(prod.split(1))
    check(skinny, "small-N contraction 4101x12x28", 1)


def test_matmul1_join_of_folds_rerolled_twice():
    """benchmarks.scala:176-187: the result's columns are separate left folds over t of A[:, t] * B[t, c] (a scalar broadcast), joined:
    the join is re-rolled into the output dimension c and every fold into a reduction over t -- one kernel, one reduction"""
    def matmul1(B, leaf, m=24, k=9, n=16):
        a, b = leaf([m, k], 1), leaf([k, n], 2)
        cols_a = a.split(1)                      # k tensors of shape [m]
        rows_b = [r.split(0) for r in b.split(0)]  # rows_b[t][c]: scalars
        outs = []
        for c in range(n):
            outs.append(chain([cols_a[t] * rows_b[t][c].broadcast([m]) for t in range(k)]))
        return B.join(outs)                      # [m, n]
    check(matmul1, "join re-rolled into an output dimension", 1)
    check(lambda B, leaf: matmul1(B, leaf, m=40, k=12, n=6), "join re-rolled into an output dimension", 1)  # n % 4 != 0: scalar lanes


def test_whole_tensor_folds_and_iterated_maps():
    check(lambda B, leaf: (leaf([33, 20], 1) * leaf([33, 20], 2)).sum(), "whole-tensor fold", 4)
    check(lambda B, leaf: B.abs(leaf([4099], 3)).sum(), "whole-tensor fold", 4)

    def iterated(B, leaf):
        x, b = leaf([8, 12], 1), leaf([8, 12], 2)
        for _ in range(12):
            x = B.max(x - b, -x)
        return x
    check(iterated, "int it_", 0)


def test_joins_at_every_dimension_and_tuple_stores():
    for d in (0, 1, 2):
        check(lambda B, leaf, d=d: B.join([B.abs(leaf([4, 8], 1)), leaf([4, 8], 2) * leaf([4, 8], 3), B.fill(2.0, [4, 8])], d), None, 0)
        check(lambda B, leaf, d=d: B.join(leaf([5, 4, 8], 4).split(1), d), None, None)


# ---- random graphs: the differential fuzzer's generator, run on the emulator -------------------------------------------------------
from test_fuzz_differential import DIMS_BIG, Gen, Pair  # noqa: E402


class EmuGen(Gen):
    """single-kernel graphs only: leaves are positions in the argument list (T.random on the cuda side, integer data on the oracle's),
    no evaluation barriers (nonInline / reshape / inner sums) inside the graph"""

    def leaf(self, shape=None):
        shape = self.shape() if shape is None else list(shape)
        pad = float(self.choice((0.0, 0.0, 3.0, -2.0)))
        kind = self.rng.randint(3)
        self.note(f"leaf{shape} pad={pad} kind={kind}")
        if kind == 1 or not shape:
            v = float(self.rng.randint(-3, 4))
            if not shape:
                return Pair(self.T.scalar(v, padding=pad), self.R.scalar(v, padding=pad), 4)
            return Pair(self.T.fill(v, shape, padding=pad), self.R.fill(v, shape, padding=pad), 4)
        data = self.rng.randint(-4, 5, size=int(np.prod(shape))).astype(np.float32).reshape(shape)
        return Pair(self.T.random(shape, seed=1, padding=pad), self.R(data, padding=pad), 4)

    def view(self, p):
        for _ in range(8):
            before = len(self.trace)
            q = super().view(p)
            if not (self.trace[before].startswith("nonInline") or self.trace[before].startswith("reshape")):
                return q
            del self.trace[before:]
        return p

    def fold(self, p):
        for _ in range(8):
            before = len(self.trace)
            q = super().fold(p)
            if len(self.trace) == before or self.trace[before] != "sum":
                return q
            del self.trace[before:]
        return p


def _emulated_fuzz(seeds, **kw):
    ran, kinds = 0, {}
    for seed in seeds:
        gen = EmuGen(cuda, seed, **kw)
        p = gen.expr(depth=2)
        if isinstance(p.r, (ref._Fill,)) or int(np.prod(p.shape)) > 20000:
            continue
        if any(t.startswith("join") for t in gen.trace[:-1]):
            continue  # an inner join is an evaluation barrier whose buffer layout differs between the backends (one-kernel join at a dimension)
        params = oracle_params(p.r, root_is_join=gen.trace[-1].startswith("join"))
        try:
            k0 = p.g.compile()
        except cuda.ComputeCudaError as e:
            raise AssertionError((seed, gen.trace, str(e)[:1500])) from None
        if k0.info.kind == 2 or k0.info.n_launches == 0:
            continue
        ords = arg_ordinals(cuda, k0)
        if len(params) == 1 and ords and min(ords) >= 1:  # views of one unevaluated inline tensor, composed into the kernel (see check())
            inner = ref.parameter_descendants(params[0].id.closure())
            if max(ords) > len(inner):
                continue
            leaves = [inner[o - 1].id.buffer() for o in ords]
        elif sorted(ords) == list(range(len(params))):
            leaves = [params[o].id.buffer() for o in ords]
        else:
            continue  # a mix of evaluated and composed parameters: the two backends number them differently
        try:
            got, k = emulate(cuda, p.g, leaves, max_threads=1 << 18)
        except AssertionError as e:
            if "too large" in str(e):
                continue
            raise AssertionError((seed, gen.trace, str(e))) from None
        want = p.r.flat_array()
        ok = tuple(p.g.shape) == tuple(p.r.shape) and (np.array_equal(got.view(np.uint32), want.view(np.uint32)) or np.array_equal(got, want))
        assert ok, (seed, gen.trace, got[:8].tolist(), want[:8].tolist())
        ran += 1
        kinds[int(k.info.kind)] = kinds.get(int(k.info.kind), 0) + 1
    return ran, kinds


@pytest.mark.parametrize("block", range(3))
def test_random_graphs_on_the_emulator(block):
    ran, kinds = _emulated_fuzz([300000 + 100 * block + case for case in range(14)])
    assert ran >= 8, (ran, kinds)


@pytest.mark.parametrize("block", range(2))
def test_larger_random_graphs_on_the_emulator(block):
    ran, kinds = _emulated_fuzz([310000 + 100 * block + case for case in range(10)], dims=DIMS_BIG, max_rank=3)
    assert ran >= 5, (ran, kinds)

