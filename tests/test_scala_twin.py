"""CPU-only: pins the tree blob a JVM front end must write. No JVM exists here, so scala/cuda/.../CudaTreeWriter.scala cannot run; its
line-by-line Python twin (tests/scala_twin.py), driven by the oracle's restatement of Tensors.scala / Trees.scala, must produce for every
spec case (tests/golden/tree_blob_cases.py) the same bytes as the C++ mirror (ids normalised) and as the committed golden blobs, and the
library must see the same kernel in both (structural hash, cache hit, cc_kernel_cache_lookup = kernelCache.getIfPresent)."""
import os

import numpy as np
import pytest

from golden.make_tree_blobs import mirror_blob
from golden.tree_blob_cases import cases
from scala_twin import blob_of, normalise_ids

from compute.scala_b200 import cuda
from oracle import reference as ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tree_blobs")
NAMES = sorted(cases(ref.Tensor))


def twin_blob(what, kind) -> bytes:
    if kind == "tensor":
        return blob_of(what)
    if kind[0] == "join":
        return blob_of(list(what), join_dimension=kind[1])
    return blob_of(what, monoid=kind[1])


def test_every_case_has_a_golden_blob_and_nothing_else_is_there():
    assert sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".bin")) == NAMES


@pytest.mark.parametrize("name", NAMES)
def test_scala_writer_twin_writes_the_mirrors_bytes(name):
    what_m, kind = cases(cuda.Tensor)[name]
    what_t, _ = cases(ref.Tensor)[name]
    mirror = mirror_blob(cuda, what_m, kind)
    twin = twin_blob(what_t, kind)
    golden = open(os.path.join(GOLDEN, name + ".bin"), "rb").read()
    assert normalise_ids(mirror) == golden, "the C++ mirror's blob changed: regenerate with tests/golden/make_tree_blobs.py if intended"
    assert normalise_ids(twin) == golden  # the Scala writer's emission order and field layout
    assert twin == normalise_ids(twin)  # the Scala writer's ids ARE 1 + the first-emission ordinal


@pytest.mark.parametrize("name", NAMES)
def test_the_library_sees_one_kernel(name):
    """compile the mirror's blob, then the twin's: same structural hash, second compile is a cache hit, and the probe-only lookup
    (what CudaTensors.kernelCache.getIfPresent calls) finds it — also when the probing term does not know its output shape"""
    what_m, kind = cases(cuda.Tensor)[name]
    what_t, _ = cases(ref.Tensor)[name]
    cuda.kernel_cache_clear()
    twin = twin_blob(what_t, kind)
    assert cuda.kernel_cache_lookup(twin) is None  # nothing cached yet: the probe never compiles
    assert cuda.kernel_cache_size() == 0
    k_mirror = cuda.compile_blob(mirror_blob(cuda, what_m, kind))
    assert k_mirror.info.cache_hit == 0
    k_twin = cuda.compile_blob(twin)
    assert k_twin.info.cache_hit == 1 and k_twin.info.structural_hash == k_mirror.info.structural_hash
    assert k_twin.info.kind == k_mirror.info.kind and k_twin.source == k_mirror.source
    probe = cuda.kernel_cache_lookup(twin)
    assert probe is not None and probe.info.structural_hash == k_mirror.info.structural_hash
    if kind == "tensor":
        # TensorsSpec.scala:50-52 probes with the bare closure, which does not carry the output shape
        from scala_twin import CudaTreeWriter

        w = CudaTreeWriter()
        bare = w.finish(w.write(what_t.closure()), ())
        # (terms whose parameters carry definitions probe with the definitions attached, as compile() writes them)
        if not any(isinstance(p, ref._Inline) for p in w.parameters):
            assert cuda.kernel_cache_lookup(bare, any_out_shape=True) is not None
            if tuple(what_t.shape) != ():
                assert cuda.kernel_cache_lookup(bare, any_out_shape=False) is None
    cuda.kernel_cache_clear()


def test_structurally_equal_fills_hit_the_cache_and_different_literals_do_not():
    """TensorsSpec.scala:37-55 and TreesSpec.scala:36-91 at the boundary the Scala side binds"""
    cuda.kernel_cache_clear()
    R = ref.Tensor
    first = blob_of(R.fill(42.0, [2, 3, 5]))
    again = blob_of(R.fill(42.0, [2, 3, 5]))
    other = blob_of(R.fill(43.0, [2, 3, 5]))
    cuda.compile_blob(first)
    assert cuda.kernel_cache_lookup(again) is not None and cuda.kernel_cache_lookup(other) is None
    a, b = R.random([4, 4], seed=1), R.random([4, 4], seed=2)
    cuda.compile_blob(blob_of(a * b + a))
    assert cuda.kernel_cache_lookup(blob_of(b * a + b)) is not None  # equal modulo parameter names
    assert cuda.kernel_cache_lookup(blob_of(a * b + b)) is None  # a different sharing pattern is a different structure
    assert cuda.kernel_cache_lookup(blob_of(R.random([4, 4], seed=1, padding=1.0) * b + a)) is None  # padding is part of the key
    cuda.kernel_cache_clear()


def test_deep_chains_do_not_recurse():
    """a 16384-term per-axis sum: the writer's explicit stack (the reference's structural hash recurses to the chain depth)"""
    R = ref.Tensor
    x = R.random([4096, 4], seed=1)
    parts = x.split(0)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    blob = blob_of(acc)
    k = cuda.compile_blob(blob)
    assert k.info.kind == 1  # re-rolled into an axis reduction
