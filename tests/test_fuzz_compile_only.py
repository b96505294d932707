"""CPU-only: the random expression graphs of tests/test_fuzz_differential.py (same generator, lazily generated leaves, extents up to
256 so that the vector, tiled-transpose, stencil-tile, re-rolled and tuple-store templates are all reached) must every one of them
plan and compile for sm_100a through NVRTC — no generator exception, no compiler error.  The values are checked on the GPU by the
differential test; this one guards the code generator where no GPU exists."""
import numpy as np
import pytest

from test_fuzz_differential import Gen, Pair

from compute.scala_b200 import cuda

DIMS_WIDE = (1, 4, 8, 12, 64, 128, 132, 256)


class LazyGen(Gen):
    """leaves that need no device: Tensor.random / fill / scalar (nothing is evaluated)"""

    def leaf(self, shape=None):
        shape = self.shape() if shape is None else list(shape)
        pad = float(self.choice((0.0, 0.0, 3.0, -2.0)))
        kind = self.rng.randint(3)
        seed = int(self.rng.randint(1, 1000))
        self.note(f"leaf{shape} pad={pad} kind={kind}")
        if kind == 1 or not shape:
            v = float(self.rng.randint(-3, 4))
            if not shape:
                return Pair(self.T.scalar(v, padding=pad), self.R.scalar(v, padding=pad), 4)
            return Pair(self.T.fill(v, shape, padding=pad), self.R.fill(v, shape, padding=pad), 4)
        return Pair(self.T.random(shape, seed=seed, padding=pad), self.R.fill(0.0, shape, padding=pad), 4)


class ShapeOnly:
    """stands in for the oracle's Tensor class: the generator only needs shapes from its second backend here"""

    Tensor = None


@pytest.mark.parametrize("block", range(6))
def test_random_graphs_plan_and_compile(block):
    from oracle import reference as ref

    kinds = {}
    for case in range(25):
        seed = 90000 + 100 * block + case
        gen = LazyGen(cuda, seed, dims=DIMS_WIDE if case % 2 else (1, 2, 3, 4, 5, 8, 12), max_rank=2 if case % 2 else 4)
        try:
            p = gen.expr(depth=2)
            if int(np.prod(p.shape)) > 2_000_000:
                continue
            k = p.g.compile()
            kinds[int(k.info.kind)] = kinds.get(int(k.info.kind), 0) + 1
            assert tuple(p.g.shape) == tuple(p.r.shape), gen.trace
        except cuda.ComputeCudaError as e:
            raise AssertionError((seed, gen.trace, str(e)[:2000])) from None
    assert sum(kinds.values()) >= 15, kinds


def test_random_windows_and_permutes_compile():
    """the templates the graph generator rarely reaches: stencil tiles (random windows, paddings, leading offsets, extra operands),
    tiled transposes of random permutations, shifted aligned-vector loads on narrow tensors"""
    T = cuda.Tensor
    rng = np.random.RandomState(7)
    seen = {"stencil tile": 0, "tiled transpose": 0, "float A": 0}
    for case in range(60):
        rank = int(rng.randint(2, 5))
        wide = case % 3 != 2
        shape = [int(rng.choice((1, 2, 3, 5))) for _ in range(rank - 2)] + [int(rng.choice((8, 17, 40, 64))), int(rng.choice((128, 132, 256, 512) if wide else (16, 32, 64)))]
        x = T.random(shape, seed=int(rng.randint(1, 99)), padding=float(rng.choice((0.0, -2.0, 7.5))))
        if case % 2 == 0:
            n = int(rng.randint(6, 26))
            lead = [int(rng.randint(-1, 2)) for _ in range(rank - 2)] if rng.rand() < 0.3 else [0] * (rank - 2)
            offs = {(int(rng.randint(-3, 4)), int(rng.randint(-6, 7))) for _ in range(n)}
            terms = [x.translate(lead + [dy, dx]) for dy, dx in sorted(offs)]
            f = (lambda a, b: a + b) if rng.rand() < 0.5 else T.max
            acc = terms[0]
            for t in terms[1:]:
                acc = f(acc, t)
            if rng.rand() < 0.5:
                acc = acc * T.random(shape, seed=5) - T.random(shape, seed=6).translate([0] * (rank - 1) + [1])
            e = acc
        else:
            perm = [int(v) for v in rng.permutation(rank)]
            e = x.permute(perm)
            if rng.rand() < 0.5:
                e = T.abs(e) + T.random(list(e.shape), seed=3)
        try:
            src = e.compile().source
        except cuda.ComputeCudaError as err:
            raise AssertionError((case, shape, str(err)[:2000])) from None
        tail = src[max(0, src.rindex('extern "C"') - 6000):]
        for tag in seen:
            if tag in tail:
                seen[tag] += 1
    assert seen["stencil tile"] >= 8 and seen["tiled transpose"] >= 3 and seen["float A"] >= 3, seen
