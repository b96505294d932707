"""The expression graphs whose tree blobs are pinned under tests/golden/tree_blobs/: the reference's own spec cases (TensorsSpec.scala,
cpuSpec.scala, cpu.scala's scaladoc examples) and the BASELINE configurations in miniature, written once against a backend `B`
(compute.scala_b200.cuda.Tensor or oracle.reference.Tensor — same API). Leaves are `random` / `fill`, which need no device.
Each case returns (what, kind) with kind one of: "tensor" (an InlineTensor's closure), ("join", dimension | None), ("reduce", monoid)."""


def fold(parts, op=lambda a, b: a + b):
    acc = parts[0]
    for p in parts[1:]:
        acc = op(acc, p)
    return acc


def matmul2(m1, m2):  # TensorsSpec.scala:472-479, benchmarks.scala:188-191
    i, j = m1.shape
    _, k = m2.shape
    product = m1.broadcast([i, j, k]) * m2.reshape([1, j, k]).broadcast([i, j, k])
    return fold(product.split(1))


def matmul1_columns(m1, m2):  # TensorsSpec.scala:506-518, benchmarks.scala:176-187
    columns1 = m1.split(1)
    return [fold([l * r.broadcast(l.shape) for l, r in zip(columns1, column2.split(0))]) for column2 in m2.split(1)]


def convolute_outputs(B, inp, weight, bias):  # benchmarks.scala:463-556, TensorsSpec.scala:144-249
    batch, height, width, depth = inp.shape
    kh, kw, _, filters = weight.shape
    input_seq, bias_seq = inp.split(3), bias.split(0)
    outs = []
    for f, khkwd in enumerate(weight.split(3)):
        terms = []
        for oy, kwd in zip(range(-(kh // 2), kh // 2 + 1), khkwd.split(0)):
            for ox, d in zip(range(-(kw // 2), kw // 2 + 1), kwd.split(0)):
                for in_c, w_c in zip(input_seq, d.split(0)):
                    terms.append(in_c.translate([0, oy, ox]) * w_c.broadcast([batch, height, width]))
        outs.append(bias_seq[f].broadcast([batch, height, width]) + fold(terms))
    return outs


def cases(B):
    r = lambda shape, seed, padding=0.0: B.random(shape, seed=seed, padding=padding)  # noqa: E731
    out = {}
    out["fill_2x3x5"] = (B.fill(42.0, [2, 3, 5]), "tensor")  # TensorsSpec.scala:37-55
    out["translate_padding_99"] = (B.fill(42.0, [2, 3, 5], padding=99.0).translate([1, 2, -3]), "tensor")  # :75-113
    out["translate_of_data"] = (r([2, 3, 5], 1, 99.0).translate([1, 2, -3]), "tensor")
    out["split_last_dimension"] = (r([1, 1, 1, 2], 2).split(3)[1], "tensor")  # :115-121
    t = r([1, 1, 2], 3)
    t2 = t + t
    out["plus_and_multiplication_shared_operand"] = (t2 * t2, "tensor")  # :131-138
    a, b, c = r([8, 12], 1), r([8, 12], 2), r([8, 12], 3)
    out["c1_tanh_a_times_b_plus_c"] = (B.tanh(a * b + c), "tensor")  # BASELINE config 1
    out["c2_chain"] = (B.tanh(B.log(B.exp(a * b + c) + a) * b) + c, "tensor")  # BASELINE config 2
    out["every_operator"] = (B.min(B.abs(a) / B.sqrt(b), B.max(-a % b, c - a)), "tensor")
    out["transpose_3d"] = (r([2, 2, 3], 4).transpose(), "tensor")  # :452-466
    out["c4_permute_translate"] = (r([4, 5, 6], 7).permute([2, 0, 1]).translate([3, -5, 7]), "tensor")  # BASELINE config 4, SURVEY A.4
    out["broadcast_2x3_to_2x3x4"] = (r([2, 3], 5).broadcast([2, 3, 4]), "tensor")  # :491-500
    out["scale_non_integer_coefficients"] = (r([3, 5], 6).scale([7, 2]), "tensor")  # Tensors.scala:950-965
    out["matmul2_2x3_3x4"] = (matmul2(r([2, 3], 8), r([3, 4], 9)), "tensor")  # :468-489 — the inline product travels as a definition
    out["matmul1_join_of_folds"] = (matmul1_columns(r([2, 3], 8), r([3, 4], 9)), ("join", None))  # :502-528
    out["axis0_sum_chain"] = (fold(r([16, 6], 5).split(0)), "tensor")  # README.md:301-310, BASELINE config 3
    out["axis1_max_chain"] = (fold(r([6, 16], 5).split(1), B.max), "tensor")
    f1, f2 = B.fill(42.0, [3, 4]), B.fill(43.0, [3, 4])
    for d in (0, 1, 2):
        out[f"join_fills_at_{d}"] = ([f1, f2], ("join", d))  # cpuSpec.scala:18-38
    x = r([2, 3, 4], 10)
    out["join_of_split_round_trip"] = (x.split(1), ("join", None))  # cpu.scala:62-93
    ni = (B.fill(2.0, [2, 3]).nonInline() if hasattr(B.fill(2.0, [2, 3]), "nonInline") else B.fill(2.0, [2, 3]).non_inline())
    out["non_inline_chain"] = (ni + ni, "tensor")  # cpuSpec.scala:9-16
    out["sum_of_inline_chain"] = (a * b + c, ("reduce", "+"))  # Tensors.scala:673-771 with the operand fused (kind 30)
    out["min_of_view"] = (r([4, 5, 6], 7).permute([2, 0, 1]), ("reduce", "min"))
    inp, weight, bias = r([2, 5, 6, 2], 1), r([3, 3, 2, 3], 2), r([3], 3)
    out["convolution_3x3"] = (convolute_outputs(B, inp, weight, bias), ("join", None))  # benchmarks.scala:463-556
    return out
