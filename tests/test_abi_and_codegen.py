"""CPU-only checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol include/compute_cuda.h
declares, the product path fails loudly without a driver, and the code generator (tree blob -> plan -> NVRTC for
sm_100a) behaves as the reference's kernel cache / Trees laws require (TreesSpec.scala:36-91, TensorsSpec.scala:37-55)."""
import ctypes as C
import struct

import numpy as np
import pytest

from compute.scala_b200 import _lib, cuda


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 70
    L = _lib.lib()
    for n in names:
        assert getattr(L, n) is not None
    assert b"sm_100a" in L.cc_version()


def test_no_cpu_fallback_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cuda.ComputeCudaError) as e:
        cuda.init()
    assert e.value.status in (-3, -4, -8)
    # compute entry points refuse to run without a context
    with pytest.raises(cuda.ComputeCudaError):
        cuda.Tensor.fill(1.0, [4]).flatArray()
    with pytest.raises(cuda.ComputeCudaError):
        cuda.Buffer.alloc(16)


def _build_c_consumer(tmp_path):
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "compute", "scala_b200")
    exe = str(tmp_path / "c1_from_c")
    cmd = ["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(root, "include"),
           os.path.join(root, "examples", "c1_from_c.c"), "-o", exe, "-L", libdir, "-lcompute_cuda", f"-Wl,-rpath,{libdir}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_a_c_caller_fails_loudly_without_a_gpu(tmp_path):
    """include/compute_cuda.h compiles as strict C11 (scalars and pointers only) and links against the library; the C program
    that drives BASELINE config 1 through it refuses to run without a driver instead of falling back to the CPU"""
    import subprocess

    import torch

    exe = _build_c_consumer(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is tests/test_threads_and_events.py::test_c_consumer_runs_config_1")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr and b"sm_100a".decode() in r.stdout


def test_product_does_not_import_the_oracle():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "compute")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "oracle_cpu" not in text, f


T = cuda.Tensor


def rnd(shape, seed=1):
    return T.random(shape, seed=seed)  # lazy: no device work until a slow action


def test_kernel_cache_is_structural():
    # same structure, different parameter identities and seeds -> one compile (TreesSpec.scala:47-66, TensorsSpec.scala:50-52)
    k1 = T.tanh(rnd([64, 64], 1) * rnd([64, 64], 2) + rnd([64, 64], 3)).compile()
    before = cuda.stats()
    k2 = T.tanh(rnd([64, 64], 7) * rnd([64, 64], 8) + rnd([64, 64], 9)).compile()
    after = cuda.stats()
    assert k1.handle == k2.handle and k2.info.cache_hit == 1
    assert after["compiles"] == before["compiles"] and after["cache_hits"] == before["cache_hits"] + 1
    # different literal / shape / padding / operand order -> different kernels (TreesSpec.scala:68-91)
    base = (T.fill(2.0, [8]) + rnd([8])).compile()
    assert (T.fill(3.0, [8]) + rnd([8])).compile().handle != base.handle
    assert (T.fill(2.0, [9]) + rnd([9])).compile().handle != base.handle
    assert (rnd([8]) + T.fill(2.0, [8])).compile().handle != base.handle
    p0 = T.random([8], seed=1, padding=0.0).translate([1]).compile()
    p1 = T.random([8], seed=1, padding=1.0).translate([1]).compile()
    assert p0.handle != p1.handle
    # sharing is part of the structure: x*x (one parameter) is not x*y (two)
    x = rnd([8])
    assert (x * x).compile().info.n_args == 1
    assert (x * rnd([8], 2)).compile().info.n_args == 2
    assert (x * x).compile().handle != (x * rnd([8], 2)).compile().handle


def test_parameter_order_is_dfs_preorder():
    # parameterDescendants (Tensors.scala:230-251): a, b, c for tanh(a*b+c)
    k = T.tanh(rnd([4, 4], 1) * rnd([4, 4], 2) + rnd([4, 4], 3)).compile()
    ords = []
    for i in range(k.info.n_args):
        o = C.c_int32()
        cuda.check(cuda._L().cc_kernel_arg_param(k.handle, i, C.byref(o)))
        ords.append(o.value)
    assert ords == [0, 1, 2]


def test_c1_kernel_is_vectorised_and_flat():
    k = T.tanh(rnd([1024, 1024], 1) * rnd([1024, 1024], 2) + rnd([1024, 1024], 3)).compile()
    assert k.info.kind == 0 and k.info.n_args == 3
    assert k.info.algorithmic_bytes == 16 * 1024 * 1024
    src = k.source
    assert "V=4" in src and "flat=1" in src and "cc_ldg4(p0 + v * 4" in src and "cc_stg4" in src


def test_views_fold_into_integer_strides_and_drop_dead_bounds_tests():
    t = rnd([512, 512, 512], 7)
    src = t.permute([2, 0, 1]).translate([3, -5, 7]).compile().source
    # out[g0,g1,g2] = T[g1+5, g2-7, g0-3]
    assert "(int)1307133 + (int)1 * g0 + (int)262144 * g1 + (int)512 * g2" in src
    assert "i0_0 < 512" in src and "i0_2 >= 0" in src and "i0_1 >= 0" in src
    assert "i0_0 >= 0" not in src and "i0_2 < 512" not in src and "i0_1 < 512" not in src  # proven in range by interval analysis
    assert "tiled transpose" in src and "tile[0][ty + 4 * r][tx]" in src  # source-contiguous along g0, output along g2
    # permute / broadcast / split can never leave the source: no tests at all, and the contiguous case vectorises
    src = rnd([512, 512], 8).reshape([1, 512, 512]).broadcast([512, 512, 512]).compile().source
    assert "?" not in src.split("// elementwise")[1].split("st(")[0]
    assert "cc_ldg4" in src
    src = rnd([512, 512], 8).broadcast([512, 512, 512]).compile().source
    assert "L0[1] = L0[2] = L0[3] = L0[0]" in src  # trailing broadcast: one scalar load feeds 4 lanes


def axis_sum(x, axis):
    parts = x.split(axis)
    acc = parts[0]
    for p in parts[1:]:
        acc = acc + p
    return acc


def test_plus_chain_is_rerolled_into_a_reduction():
    x = rnd([256, 512])
    k0 = axis_sum(x, 0).compile()
    assert k0.info.kind == 1 and k0.info.n_args == 1 and "column owner" in k0.source
    assert k0.info.algorithmic_bytes == 4 * 256 * 512 + 4 * 512
    k1 = axis_sum(x, 1).compile()
    assert k1.info.kind == 1 and "row owner" in k1.source
    # short chains stay unrolled in one elementwise kernel, as the reference runs them
    assert axis_sum(rnd([4, 16]), 0).compile().info.kind == 0
    # a chain whose terms are not congruent is not a reduction
    y = rnd([8, 16])
    parts = y.split(0)
    acc = parts[0]
    for i, p in enumerate(parts[1:]):
        acc = acc + (p * p if i == 3 else p)
    assert acc.compile().info.kind == 0


def matmul2(a, b):
    i, j = a.shape
    _, k = b.shape
    product = a.broadcast([i, j, k]) * b.reshape([1, j, k]).broadcast([i, j, k])
    return axis_sum(product, 1)


def test_min_max_times_chains_are_rerolled_like_plus_chains():
    """MonoidPrograms is generic over append / zero (Tensors.scala:308-311): `t.split(axis).reduce(Tensor.max)` etc. are the same
    idiom as the per-axis sum and must not become a 4096-term unrolled kernel"""
    x = rnd([4096, 64])

    def chain(parts, f):
        acc = parts[0]
        for p in parts[1:]:
            acc = f(acc, p)
        return acc

    for f, name in ((T.max, "Max"), (T.min, "Min"), (lambda a, b: a * b, "Times")):
        for axis, owner in ((0, "column owner"), (1, "row owner")):
            k = chain(x.split(axis), f).compile()
            assert k.info.kind == 1 and owner in k.source and f"fold={name}" in k.source, (name, axis)
    # an epilogue around a max chain (the softmax numerator's shift) and a max chain that never becomes a contraction
    m = chain(x.split(1), T.max)
    k = (T.exp(m) - T.fill(1.0, [4096])).compile()
    assert k.info.kind == 1 and "epilogue=1" in k.source and "fold=Max" in k.source
    a, b = rnd([256, 512], 1), rnd([512, 256], 2)
    prod = a.broadcast([256, 512, 256]) * b.reshape([1, 512, 256]).broadcast([256, 512, 256])
    k = chain(prod.split(1), T.max).compile()  # max_t a[i,t] * b[t,k]: "tropical" product, generic reduction, not tcgen05
    assert k.info.kind == 1 and "fold=Max" in k.source
    # any tree of one operator folds its leaves left to right: pairwise (parallel-collection) reduce, reduceRight
    def pairwise(parts, f):
        while len(parts) > 1:
            parts = [f(parts[i], parts[i + 1]) if i + 1 < len(parts) else parts[i] for i in range(0, len(parts), 2)]
        return parts[0]

    def right(parts, f):
        acc = parts[-1]
        for q in reversed(parts[:-1]):
            acc = f(q, acc)
        return acc

    import time

    t0 = time.perf_counter()
    for red in (pairwise, right):
        for f, name in ((lambda a, b: a + b, "Plus"), (T.min, "Min")):
            k = red(x.split(0), f).compile()  # 4096 leaves
            assert k.info.kind == 1 and f"fold={name}" in k.source and "T=4096" in k.source
    assert time.perf_counter() - t0 < 20.0  # linear, not quadratic, in the number of leaves
    shared = x.split(0)[0] + x.split(0)[1]
    assert pairwise([shared, shared] + x.split(0)[2:12], lambda a, b: a + b).compile().info.kind == 0  # a shared partial sum is a leaf: no step
    # mixed operators do not form one chain
    parts = x.split(1)
    mixed = T.max(T.max(parts[0] + parts[1], parts[2]), parts[3])
    assert mixed.compile().info.kind == 0


def test_matmul_pattern_is_recognised_through_the_fusion_barrier():
    k = matmul2(rnd([64, 48], 1), rnd([48, 32], 2)).compile()
    info = k.info
    assert info.kind in (1, 2)
    assert info.n_args == 2  # A and B — never the i*j*k product
    assert "inline operand composed into the reduction" in k.source
    big = matmul2(rnd([256, 128], 1), rnd([128, 256], 2)).compile()
    assert big.info.flops in (2 * 256 * 128 * 256, 2 * 256 * 128 * 256)


def matmul1(m1, m2):
    cols1 = m1.split(1)
    outs = []
    for col2 in m2.split(1):
        terms = [l * r.broadcast(l.shape) for l, r in zip(cols1, col2.split(0))]
        acc = terms[0]
        for x in terms[1:]:
            acc = acc + x
        outs.append(acc)
    return T.join(outs)


def test_matmul1_join_of_folds_is_rerolled_twice():
    # benchmarks.scala:176-187: Concatenate over c of left folds over t -> one output dimension + one reduction index
    k = matmul1(rnd([4096, 32], 1), rnd([32, 32], 2)).compile()
    assert k.info.kind == 1 and k.info.n_args == 2 and "join re-rolled into an output dimension; Plus chain of 32 congruent terms re-rolled into a reduction over 32" in k.source
    assert k.info.algorithmic_bytes == 4 * (4096 * 32 + 32 * 32 + 4096 * 32)
    big = matmul1(rnd([1024, 128], 1), rnd([128, 256], 2)).compile()
    assert big.info.kind == 2 and big.info.flops == 2 * 1024 * 128 * 256
    # a join whose chains differ in more than the affine step stays a tuple store of unrolled chains
    a, b = rnd([64, 8], 1), rnd([8, 2], 2)
    cols = a.split(1)
    c0 = [l * r.broadcast(l.shape) for l, r in zip(cols, b.split(1)[0].split(0))]
    c1 = [l + r.broadcast(l.shape) for l, r in zip(cols, b.split(1)[1].split(0))]
    f = lambda ts: __import__("functools").reduce(lambda x, y: x + y, ts)
    assert T.join([f(c0), f(c1)]).compile().info.kind == 0


def test_contraction_accepts_any_shape_and_skips_tiny_products():
    assert matmul2(rnd([1000, 300], 1), rnd([300, 700], 2)).compile().info.kind == 2
    assert matmul2(rnd([1029, 33], 1), rnd([33, 2000], 2)).compile().info.kind == 2
    assert matmul2(rnd([48, 40], 1), rnd([40, 24], 2)).compile().info.kind == 1      # tiny: one generic launch beats three
    assert matmul2(rnd([65536, 32], 1), rnd([32, 16], 2)).compile().info.kind == 1   # skinny N: HBM-bound, no workspace traffic


def test_whole_tensor_fold_fuses_the_operand():
    a, b, c = rnd([1024, 1024], 1), rnd([1024, 1024], 2), rnd([1024, 1024], 3)
    k = T.tanh(a * b + c).sum().compile()
    assert k.info.kind == 4 and k.info.n_args == 3 and k.info.n_launches == 1
    assert k.info.algorithmic_bytes == 12 * 1024 * 1024 + 4
    src = k.source
    assert "monoid=cc_plus V=4 U=4 flat=1" in src and "cc_fold_finish<M>" in src and "cc_tanh" in src
    assert "monoid=cc_max" in rnd([5, 7, 3]).translate([1, 0, -1]).reduce("max").compile().source
    assert "monoid=cc_times" in rnd([8]).product().compile().source
    with pytest.raises(ValueError):
        rnd([8]).reduce("/")
    # Reduce is a root-only node
    L = cuda._L()
    h = C.c_uint64()
    blob = struct.pack("<4I", 0x31544343, 3, 2, 0) + struct.pack("<If", 1, 1.0) + struct.pack("<4I", 30, 22, 0, 0) + struct.pack("<2I", 15, 1)
    assert L.cc_compile(blob, len(blob), C.byref(h)) == -6


def convolute(inp, weight, bias):
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
        acc = summands[0]
        for x in summands[1:]:
            acc = acc + x
        outs.append(bias_seq[f].broadcast([batch, height, width]) + acc)
    return T.join(outs)


def test_convolution_is_a_nested_reduction_with_an_epilogue():
    # 3 x 3 x depth terms per output channel, affine in (kernel row, kernel column, channel); `bias + chain` is the epilogue
    k = convolute(rnd([2, 32, 32, 8], 1), rnd([3, 3, 8, 8], 2), rnd([8], 3)).compile()
    assert k.info.kind == 1 and k.info.n_args == 3
    src = k.source
    assert "Plus chain of 72 congruent terms re-rolled into a reduction over 3 x 3 x 8 with an elementwise epilogue" in src
    assert "T=3x3x8" in src and "epilogue=1" in src and "post(acc" in src
    assert "cc_ldc4(p1" in src  # the weights are reused by every output pixel: L1-cached loads
    # padding tests survive only where the 3x3 window can leave the image (rows / columns), never on batch or channel
    assert "i0_1 >= 0 && i0_1 < 32 && i0_2 >= 0 && i0_2 < 32" in src and "i0_0" not in src and "i0_3" not in src
    # many pixels, few filters, short K (the reference's own benchmark sizes, benchmarks.scala:612-630): one kernel on warp-level MMAs with the
    # weights in registers; the bias epilogue is applied to the accumulator fragments
    small = convolute(rnd([128, 32, 32, 8], 1), rnd([3, 3, 8, 8], 2), rnd([8], 3)).compile()
    assert small.info.kind == 1 and small.info.n_launches == 1 and small.info.flops == 2 * 128 * 32 * 32 * 8 * 72
    assert "small-N contraction 131072x8x72 on warp-level MMAs" in small.source and "cc_mma_tf32_16x8x8" in small.source and "post1(" in small.source
    # large enough (and enough filters to reuse each gathered row): an implicit GEMM over gathered operand panels
    big = convolute(rnd([64, 56, 56, 64], 1), rnd([3, 3, 64, 64], 2), rnd([64], 3)).compile()
    assert big.info.kind == 2 and big.info.n_launches == 4 and big.info.flops == 2 * 64 * 56 * 56 * 64 * 576
    assert "general contraction 200704x64x576 over gathered operand panels" in big.source
    assert "panel_a" in big.source and "panel_b" in big.source and "post_kernel" in big.source and "cc_split_tf32" in big.source
    # an epilogue around a plain per-axis sum
    x, b = rnd([64, 512], 1), rnd([512], 2)
    e = T.tanh(axis_sum(x, 0) + b)
    ke = e.compile()
    assert ke.info.kind == 1 and "with an elementwise epilogue" in ke.source and "cc_tanh" in ke.source


def test_join_is_rerolled_into_an_output_dimension():
    t = rnd([16, 8, 32])
    k = T.join(t.split(1)).compile()
    assert k.info.kind == 0 and "join re-rolled" in k.source and k.info.out_floats == 16 * 8 * 32
    k2 = T.join([rnd([4, 4], 1), rnd([4, 4], 2) * rnd([4, 4], 3)]).compile()
    assert "per-index tuple stores" in k2.source and k2.info.out_floats == 32


def test_bad_blobs_are_rejected():
    L = cuda._L()
    h = C.c_uint64()
    for blob in (b"", b"\x00" * 16, struct.pack("<4I", 0x31544343, 1, 5, 0), struct.pack("<4I", 0x31544343, 1, 0, 0) + struct.pack("<I", 99)):
        st = L.cc_compile(blob, len(blob), C.byref(h))
        assert st == -6, st
        assert L.cc_last_error()


def test_error_behaviour_matches_the_reference():
    with pytest.raises(ValueError):  # IllegalArgumentException, Tensors.scala:208-222
        rnd([2, 3]) + rnd([3, 3])
    with pytest.raises(ValueError):  # Tensors.scala:879-882
        rnd([2, 3]).reshape([4, 2])
    with pytest.raises(ValueError):  # Tensors.scala:1009-1011
        rnd([2, 3]).permute([0])
    with pytest.raises(ValueError):  # Tensors.scala:971-973
        rnd([2, 3]).translate([1.0])
    with pytest.raises(ValueError):  # Tensors.scala:845-848
        rnd([2, 3]).broadcast([2, 4])
    with pytest.raises(ValueError):
        T([[1.0], [2.0, 3.0]])
    assert (rnd([2, 1]) + rnd([1, 3])).shape == (2, 3)
    assert (rnd([2]) + rnd([2, 3])).shape == (2, 3)
    assert rnd([2, 3, 4]).transpose().shape == (4, 3, 2)
    assert [p.shape for p in rnd([2, 3, 4]).split(1)] == [(2, 4)] * 3
    assert T.join(rnd([2, 3, 4]).split(1), 1).shape == (2, 3, 4)


def test_deep_chains_do_not_overflow_the_stack():
    x = rnd([16384, 8])
    k = axis_sum(x, 0).compile()  # 16384-term Plus chain (the JVM recurses to this depth, Trees.scala:70-91)
    assert k.info.kind == 1
    del x, k
    assert cuda.live_tensors() >= 0


def test_iterated_maps_become_counted_loops():
    """`(0 until n).foldLeft(x)(f)` (benchmarks.scala:100-108, 319-326): n copies of f in the tree, one loop in the kernel"""
    import time

    def fold(n, x, f):
        for _ in range(n):
            x = f(x)
        return x

    t0 = time.perf_counter()
    k = fold(100, rnd([128, 128]), T.tanh).compile()
    assert time.perf_counter() - t0 < 1.0  # 100 inlined tanhf bodies took NVRTC 1.4-1.8 s
    assert "for (int it_ = 0; it_ < 100; ++it_)" in k.source and k.source.count("= cc_tanh(") == 1
    a, b, c = rnd([32, 32], 1), rnd([32, 32], 2), rnd([32, 32], 3)
    k = fold(100, a, lambda v: v * b + c).compile()
    assert "it_ < 99" in k.source and k.info.n_args == 3  # the first a*b+c interleaves the loads of b and c: 99 uniform periods
    # intermediate values that are read again later keep the chain unrolled ...
    steps = [a]
    for _ in range(10):
        steps.append(T.tanh(steps[-1]))
    assert "int it_" not in (steps[-1] + steps[5]).compile().source
    # ... but the last value of a loop may be used as often as needed, and short chains are left alone
    e = fold(12, a, T.tanh)
    assert "it_ < 12" in (e * e + e).compile().source
    assert "int it_" not in fold(7, a, T.tanh).compile().source
    # a period of several ops with two carried values: (p, q) -> (p + q, p * q) would need tuples; x -> exp(x) * x + b does not
    k = fold(9, a, lambda v: T.exp(v) * v + b).compile()
    assert "it_ < 8" in k.source or "it_ < 9" in k.source
    # inside a re-rolled reduction's term
    x = rnd([64, 256])
    k = axis_sum(T.tanh(T.tanh(T.tanh(T.tanh(T.tanh(T.tanh(T.tanh(T.tanh(x)))))))).nonInline(), 0).compile()
    assert k.info.kind == 1


def test_tuned_template_choices_are_stable():
    """the template decisions the B200 measurements settled (profiles/r01_view_sweep.json), pinned through the generated source"""
    def gen(e):
        src = e.compile().source
        return src[src.rindex("\n// ", 0, src.rindex('extern "C"')):]

    x = rnd([512, 512])
    # a source read through several views: L1-allocating loads; a single translated view keeps streaming (no_allocate) loads
    two = gen(x + x.translate([1, 0]))
    assert "cc_ldc4(" in two and "cc_ldg4(" not in two
    assert "cc_ldg4(" in gen(x.translate([1, 0]) * rnd([512, 512], 2))
    # dense windows over one source (>= 4 translated views, extents >= 8 x 128): staged through shared memory; narrower tensors stay
    # unrolled with shifted views read as aligned vector pairs and static lane picks
    five = gen(x + x.translate([0, 1]) + x.translate([0, -1]) + x.translate([1, 0]) + x.translate([-1, 0]))
    assert "elementwise" in five and "float A" not in five  # five views: L1-resident elementwise kernel, per-lane shifted loads
    terms = [x.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    acc = terms[0]
    for t in terms[1:]:
        acc = T.max(acc, t)
    win = acc.compile()
    assert win.info.kind == 0 and "stencil tile" in win.source and "R2[5], R2[6], R2[7], R2[8]" in win.source and "R5[3]" in win.source  # 4 output rows + 2 halo rows per thread
    narrow = rnd([512, 64])
    nt = [narrow.translate([dy, dx]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    acc = nt[0]
    for t in nt[1:]:
        acc = T.max(acc, t)
    nw = acc.compile()
    assert nw.info.kind == 0 and "float A" in nw.source and "& 3]" in nw.source
    # ... while a weighted window (the convolution idiom) is still a re-rolled reduction
    w = rnd([3, 3], 5)
    wt = [x.translate([dy, dx]) * w.split(0)[dy + 1].split(0)[dx + 1].broadcast([512, 512]) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    acc = wt[0]
    for t in wt[1:]:
        acc = acc + t
    assert acc.compile().info.kind == 1
    # odd fastest dimension under a view: scalar lanes, 8 elements in flight per thread
    odd = gen(rnd([1001, 1003, 127]).translate([0, 0, 1]))
    assert "V=1 U=8" in odd
    # short rows: a warp per output; long rows: a CTA per output
    def rowsum(shape):
        parts = rnd(shape).split(len(shape) - 1)
        a = parts[0]
        for q in parts[1:]:
            a = a + q
        return gen(a)
    assert "threads/output=32" in rowsum([16384, 1024]) and "threads/output=256" in rowsum([16384, 4096])
    # join(split(d), d) is the identity copy whatever d
    y = rnd([64, 96, 128])
    for d in (0, 1, 2):
        assert "flat=1" in gen(T.join(y.split(d), d))


def test_on_disk_cubin_cache(tmp_path):
    """opt-in: the cubin of a generated source survives the process; a hit skips NVRTC; a corrupt entry is recompiled"""
    import os
    import subprocess
    import sys

    d = str(tmp_path / "cubins")
    cuda.kernel_disk_cache(d)
    try:
        cuda.kernel_cache_clear()
        s0 = cuda.stats()
        k = (T.fill(1234.5, [8, 8]) * rnd([8, 8]) + rnd([8, 8], 2)).compile()
        s1 = cuda.stats()
        assert s1["nvrtc_compiles"] == s0["nvrtc_compiles"] + 1 and s1["disk_cache_hits"] == s0["disk_cache_hits"]
        files = [f for f in os.listdir(d) if f.endswith(".cubin")]
        assert len(files) == 1 and os.path.getsize(os.path.join(d, files[0])) > 1000
        # same process, in-memory cache dropped: served from disk, no NVRTC
        cuda.kernel_cache_clear()
        k2 = (T.fill(1234.5, [8, 8]) * rnd([8, 8]) + rnd([8, 8], 2)).compile()
        s2 = cuda.stats()
        assert s2["nvrtc_compiles"] == s1["nvrtc_compiles"] and s2["disk_cache_hits"] == s1["disk_cache_hits"] + 1
        assert k2.source == k.source
        # another process finds it through the environment variable
        code = (
            "from compute.scala_b200 import cuda; T = cuda.Tensor\n"
            "k = (T.fill(1234.5, [8, 8]) * T.random([8, 8], seed=1) + T.random([8, 8], seed=2)).compile()\n"
            "s = cuda.stats(); print(s['nvrtc_compiles'], s['disk_cache_hits'])"
        )
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        # (the key includes the NVRTC version: a process that imported torch first resolves torch's bundled libnvrtc, a fresh one the
        # system's, so the first fresh process may compile once more; the second one must not)
        for attempt in range(2):
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, env=dict(os.environ, CC_KERNEL_CACHE_DIR=d))
            assert r.returncode == 0, r.stderr[-500:]
        assert r.stdout.split() == ["0", "1"], (r.stdout, r.stderr[-500:])
        files = sorted(os.listdir(d), key=lambda f: os.path.getmtime(os.path.join(d, f)))[:1]  # the entry this process wrote
        # a damaged entry is ignored and rewritten
        path = os.path.join(d, files[0])
        with open(path, "r+b") as f:
            f.seek(8)
            f.write(b"\x00" * 8)
        cuda.kernel_cache_clear()
        (T.fill(1234.5, [8, 8]) * rnd([8, 8]) + rnd([8, 8], 2)).compile()
        s3 = cuda.stats()
        assert s3["nvrtc_compiles"] == s2["nvrtc_compiles"] + 1
        with open(path, "rb") as f:
            assert f.read(16)[8:] != b"\x00" * 8
    finally:
        cuda.kernel_disk_cache(None)
        cuda.kernel_cache_clear()
    before = cuda.stats()["disk_cache_hits"]
    (T.fill(1234.5, [8, 8]) * rnd([8, 8]) + rnd([8, 8], 2)).compile()
    assert cuda.stats()["disk_cache_hits"] == before  # off again


def test_kernel_cache_policy():
    """kernelCache (Tensors.scala:1267-1289): unbounded by default, an optional LRU limit, clearCache"""
    cuda.kernel_cache_clear()
    assert cuda.kernel_cache_size() == 0
    ks = [(T.fill(float(i), [8]) + rnd([8])).compile() for i in range(6)]  # literals are part of the key: 6 kernels
    assert cuda.kernel_cache_size() == 6
    cuda.kernel_cache_limit(3)
    assert cuda.kernel_cache_size() == 3
    assert ks[0].source  # evicted from the cache but the caller's handle keeps it alive
    before = cuda.stats()["compiles"]
    again = (T.fill(5.0, [8]) + rnd([8])).compile()  # most recent: still cached
    assert again.info.cache_hit == 1 and cuda.stats()["compiles"] == before
    evicted = (T.fill(0.0, [8]) + rnd([8])).compile()  # oldest: compiled again
    assert cuda.stats()["compiles"] == before + 1 and evicted.handle != ks[0].handle
    assert cuda.kernel_cache_size() == 3
    cuda.kernel_cache_limit(0)
    cuda.kernel_cache_clear()
    assert cuda.kernel_cache_size() == 0


def test_hot_calls_are_bound_natively_and_report_typed_errors(monkeypatch):
    """doBuffer / Buffer.release go through the _hotcalls CPython extension (csrc/py_hotcalls.c) when it is built and through ctypes
    otherwise; both land in libcompute_cuda.so and report the same typed errors (no GPU here: the calls must fail loudly)."""
    import importlib
    import sys

    def probe():
        e = cuda.Tensor.fill(1.0, [4, 4]) + cuda.Tensor.fill(2.0, [4, 4])
        with pytest.raises(cuda.ComputeCudaError) as ei:
            e.doBuffer()
        assert "CC_ERR_NOT_INITIALIZED" in str(ei.value) or "CC_ERR_NO_DRIVER" in str(ei.value)
        stale = cuda.Buffer(0x1234560)
        with pytest.raises(cuda.IllegalArgumentException):
            stale.release()
        assert stale.handle == 0  # a failed release is not retried by __del__

    hot = cuda._hot()
    assert type(hot.do_buffer).__name__ == "builtin_function_or_method", "the _hotcalls extension is not built (python -m compute.scala_b200.build)"
    probe()
    # the ctypes binding of the same two entry points
    import compute.scala_b200 as package

    monkeypatch.setitem(sys.modules, "compute.scala_b200._hotcalls", None)
    monkeypatch.delattr(package, "_hotcalls")
    hot = cuda._hot()
    assert type(hot.do_buffer).__name__ == "function"
    probe()
    monkeypatch.undo()
    importlib.invalidate_caches()
    assert type(cuda._hot().do_buffer).__name__ == "builtin_function_or_method"


def test_a_failing_compilation_wakes_its_waiters():
    """the JIT runs outside the runtime lock with an in-flight marker per structure: a structure whose planning fails must release the
    marker and report the error to every thread that asked for it (none may wait forever), and leave the compiler usable"""
    import threading

    L = cuda._L()
    # parses, but a Reduce root must produce exactly one float: rejected by the planner (after the marker is set)
    blob = struct.pack("<4I", 0x31544343, 2, 1, 1) + struct.pack("<i", 2) + struct.pack("<If", 1, 1.0) + struct.pack("<4I", 30, 22, 0, 0)
    results = []

    def ask():
        h = C.c_uint64()
        results.append(L.cc_compile(blob, len(blob), C.byref(h)))

    for _ in range(3):
        ts = [threading.Thread(target=ask) for _ in range(8)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=60)
        assert not any(t.is_alive() for t in ts), "a waiter was never woken"
    assert results == [-6] * 24
    k = (cuda.Tensor.random([8, 8], seed=1) + cuda.Tensor.fill(1.0, [8, 8])).compile()
    assert k.info.kind == 0


def test_every_python_operator_reaches_its_node_kind():
    """the Python view binds its operators to the C ABI's node codes directly: each must produce its own operation in the kernel text"""
    import re

    T = cuda.Tensor
    a, b = T.random([8, 8], seed=1), T.random([8, 8], seed=2)
    want = {"+": (a + b, "(_0+_1)"), "-": (a - b, "(_0-_1)"), "*": (a * b, "(_0*_1)"), "/": (a / b, "(_0/_1)"), "%": (a % b, "fmodf(_0,_1)"),
            "neg": (-a, "(-_0)"), "min": (T.min(a, b), "fminf(_0,_1)"), "max": (T.max(a, b), "fmaxf(_0,_1)"), "abs": (T.abs(a), "fabsf(_0)"),
            "sqrt": (T.sqrt(a), "sqrtf(_0)"), "tanh": (T.tanh(a), "cc_tanh(_0)"), "exp": (T.exp(a), "cc_exp(_0)"), "log": (T.log(a), "cc_log(_0)")}
    for name, (e, text) in want.items():
        lines = [l.replace(" ", "") for l in e.compile().source.split("\n") if re.match(r"\s*const float _[12] = ", l)]
        assert lines and text in lines[-1], (name, lines[-1:])
    assert (+a) is a


def test_planning_switches_are_sampled_with_the_kernel_cache_not_read_live(monkeypatch):
    """make_plan read CC_* switches through getenv while the cache is keyed by structure alone, so a switch flipped after the
    first compile changed new plans but not cached ones. They are sampled when the cache is empty (first use / cc_kernel_cache_clear):
    flipping one without clearing changes NOTHING, clearing makes it take effect for everything."""
    from compute.scala_b200 import cuda

    T = cuda.Tensor

    def column_sum(rows, cols):
        parts = T.random([rows, cols], seed=1).split(0)
        acc = parts[0]
        for p in parts[1:]:
            acc = acc + p
        return acc

    monkeypatch.delenv("CC_FUSE_COL_STAGE", raising=False)
    cuda.kernel_cache_clear()
    try:
        assert column_sum(300, 64).compile().info.n_launches == 1  # default: second stage fused
        monkeypatch.setenv("CC_FUSE_COL_STAGE", "0")
        assert column_sum(300, 64).compile().info.n_launches == 1  # cached structure: unchanged
        assert column_sum(320, 64).compile().info.n_launches == 1  # a NEW structure plans with the sampled switches too
        cuda.kernel_cache_clear()
        assert column_sum(300, 64).compile().info.n_launches == 2 and column_sum(320, 64).compile().info.n_launches == 2
    finally:
        monkeypatch.delenv("CC_FUSE_COL_STAGE", raising=False)
        cuda.kernel_cache_clear()
