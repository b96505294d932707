"""Named expression builders shared by the fixture generator (make_fixtures.py), the CPU test that pins the oracle to the
committed fixtures and the GPU test that holds the CUDA path to the same bits.  Every builder takes the Tensor class of the backend
under test (oracle `reference.Tensor` or `cuda.Tensor`: same API) and returns a lazy tensor.

bar: "exact"  bit-for-bit (index / view / copy / integer-valued work, RNG hashes)
     "ulp2"   <= 2 ulp per element (single fused elementwise ops; fma contraction allowed: compared against both oracle variants)
     "rel1e-5" <= 1e-5 relative to the largest magnitude (reductions, matmul)
"""
import numpy as np


def _chain(parts, f=lambda a, b: a + b):
    acc = parts[0]
    for p in parts[1:]:
        acc = f(acc, p)
    return acc


def _ints(T, shape, seed):  # integers in {-4..4}: every summation order is exact (dataset E)
    r = T.random(shape, seed=seed) * T.fill(9.0, shape)
    return (r - r % T.fill(1.0, shape)) - T.fill(4.0, shape)


def _matmul2(T, a, b):  # benchmarks.scala:188-191
    i, j = a.shape
    _, k = b.shape
    return _chain((a.broadcast([i, j, k]) * b.reshape([1, j, k]).broadcast([i, j, k])).split(1))


def _matmul1(T, a, b):  # benchmarks.scala:176-187
    cols = a.split(1)
    return T.join([_chain([l * r.broadcast(l.shape) for l, r in zip(cols, c2.split(0))]) for c2 in b.split(1)])


def _convolute(T, inp, weight, bias):  # TensorsSpec.scala:144-210
    batch, height, width, _ = inp.shape
    input_seq = inp.split(3)
    outs = []
    for w_f, b_f in zip(weight.split(3), bias.split(0)):
        summands = []
        for oy, w_row in zip((-1, 0, 1), w_f.split(0)):
            for ox, w_px in zip((-1, 0, 1), w_row.split(0)):
                for in_c, w_c in zip(input_seq, w_px.split(0)):
                    summands.append(in_c.translate([0, oy, ox]) * w_c.broadcast([batch, height, width]))
        outs.append(b_f.broadcast([batch, height, width]) + _chain(summands))
    return T.join(outs)


CASES = {
    # RNG (Tensors.scala:106-117, 398-443)
    "random_3x3_seed12345": (lambda T: T.random([3, 3], seed=12345), "exact"),
    "random_4099_seed_minus5": (lambda T: T.random([4099], seed=-5), "exact"),
    "random_normal_3x3x3_seed54321": (lambda T: T.randomNormal([3, 3, 3], seed=54321), "ulp2"),
    # C1 / C2 in miniature
    "c1_tanh_fma_33x65": (lambda T: T.tanh(T.random([33, 65], seed=1) * T.random([33, 65], seed=2) + T.random([33, 65], seed=3)), "ulp2"),
    "abs_sqrt_min_max_div_7x5": (
        lambda T: T.max(T.sqrt(T.random([7, 5], seed=1) + T.fill(1.0, [7, 5])) / (T.random([7, 5], seed=2) + T.fill(1.0, [7, 5])), T.abs(-T.random([7, 5], seed=3)))
        - T.min(T.random([7, 5], seed=2), T.random([7, 5], seed=3)),
        "ulp2",
    ),
    # C4: views, bit-exact
    "c4_permute_translate_9x10x11": (lambda T: T.random([9, 10, 11], seed=7).permute([2, 0, 1]).translate([3, -5, 7]), "exact"),
    "c4_broadcast_trailing_6x7_to_6x7x5": (lambda T: T.random([6, 7], seed=8).broadcast([6, 7, 5]), "exact"),
    "c4_broadcast_leading_6x7_to_5x6x7": (lambda T: T.random([6, 7], seed=8).reshape([1, 6, 7]).broadcast([5, 6, 7]), "exact"),
    "c4_split1_join_roundtrip_4x6x8": (lambda T: T.join(T.random([4, 6, 8], seed=7).split(1), 1), "exact"),
    "translate_padding_99": (lambda T: T.fill(42.0, [2, 3, 5], padding=99.0).translate([1, 2, -3]), "exact"),
    "translate_fractional_truncation": (lambda T: T.random([6, 4], seed=1).translate([0.5, -0.25]), "exact"),
    "scale_3x3_to_7x5": (lambda T: T.random([3, 3], seed=1).scale([7, 5]), "exact"),
    "transpose_2x2x3": (lambda T: T([[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [[7.0, 8.0, 9.0], [10.0, 11.0, 12.0]]]).transpose(), "exact"),
    "join_dim0_3x4": (lambda T: T.join([T.fill(42.0, [3, 4]), T.fill(43.0, [3, 4])], 0), "exact"),
    "join_dim1_3x4": (lambda T: T.join([T.fill(42.0, [3, 4]), T.fill(43.0, [3, 4])], 1), "exact"),
    # C3: reductions on exactly summable data
    "c3_full_sum_ints_64x48": (lambda T: _ints(T, [64, 48], 5).sum(), "exact"),
    "c3_axis0_sum_ints_64x48": (lambda T: _chain(_ints(T, [64, 48], 5).split(0)), "exact"),
    "c3_axis1_sum_ints_64x48": (lambda T: _chain(_ints(T, [64, 48], 5).split(1)), "exact"),
    "c3_full_sum_uniform_1000": (lambda T: T.random([1000], seed=5).sum(), "rel1e-5"),
    "sum_fill_8x8": (lambda T: T.fill(15625.0, [8, 8]).sum(), "exact"),
    "axis_max_33x20": (lambda T: _chain(T.random([33, 20], seed=6).split(0), T.max), "exact"),
    # C5: both matmul formulations
    "c5_matmul2_golden_2x3x4": (lambda T: _matmul2(T, T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), T([[7.0, 8.0, 9.0, 10.0], [11.0, 12.0, 13.0, 14.0], [15.0, 16.0, 17.0, 18.0]])), "exact"),
    "c5_matmul1_golden_2x3x4": (lambda T: _matmul1(T, T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), T([[7.0, 8.0, 9.0, 10.0], [11.0, 12.0, 13.0, 14.0], [15.0, 16.0, 17.0, 18.0]])), "exact"),
    "c5_matmul2_ints_24x40x16": (lambda T: _matmul2(T, _ints(T, [24, 40], 9), _ints(T, [40, 16], 10)), "exact"),
    "c5_matmul2_uniform_16x32x8": (lambda T: _matmul2(T, T.random([16, 32], seed=9), T.random([32, 8], seed=10)), "rel1e-5"),
    # the reference's convolution (TensorsSpec.scala:140-249) on integer-valued data
    "convolution_2x4x5x3_to_2": (lambda T: _convolute(T, _ints(T, [2, 4, 5, 3], 1), _ints(T, [3, 3, 3, 2], 2), _ints(T, [2], 3)), "exact"),
}
