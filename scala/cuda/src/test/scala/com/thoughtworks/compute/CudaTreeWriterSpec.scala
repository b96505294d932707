package com.thoughtworks.compute

import java.nio.file.{Files, Paths}
import java.nio.{ByteBuffer, ByteOrder}

import org.lwjgl.system.MemoryUtil
import org.scalatest.{FreeSpec, Matchers}

/** Holds [[CudaTreeWriter]] to the golden tree blobs of `tests/golden/tree_blobs/` (written by the C++ mirror of
  * [[CudaTensors]], `python tests/golden/make_tree_blobs.py`; the same files pin a Python transliteration of the writer in
  * `tests/test_scala_twin.py`). Each case builds the expression of `tests/golden/tree_blob_cases.py` through the public Tensor API
  * and compares the bytes `compile` would hand to `cc_compile_ex`. Needs no GPU kernels to run, only the library to load.
  */
class CudaTreeWriterSpec extends FreeSpec with Matchers {
  import cuda._

  private val goldenDirectory = Paths.get(System.getProperty("com.thoughtworks.compute.cuda.goldens", "tests/golden/tree_blobs"))

  private def golden(name: String): Array[Byte] = Files.readAllBytes(goldenDirectory.resolve(name + ".bin"))

  private def bytesOf(blob: ByteBuffer): Array[Byte] = {
    val bytes = new Array[Byte](blob.remaining())
    blob.duplicate().order(ByteOrder.LITTLE_ENDIAN).get(bytes)
    MemoryUtil.memFree(blob)
    bytes
  }

  private def fold(tensors: Seq[Tensor], operator: (Tensor, Tensor) => Tensor = _ + _): Tensor = tensors.reduce[Tensor](operator)

  private def r(shape: Array[Int], seed: Int, padding: Float = 0.0f): Tensor = Tensor.random(shape, seed = seed, padding = padding)

  /** name -> the blob the backend writes for it ([[CudaTensors.treeBlobOf]] is the test hook onto `compile`'s writer) */
  private val cases: Seq[(String, () => ByteBuffer)] = {
    val (a, b, c) = (r(Array(8, 12), 1), r(Array(8, 12), 2), r(Array(8, 12), 3))
    val t = r(Array(1, 1, 2), 3)
    val doubled = t + t
    def matmul2(m1: Tensor, m2: Tensor): Tensor = {
      val Array(i, j) = m1.shape
      val Array(_, k) = m2.shape
      fold((m1.broadcast(Array(i, j, k)) * m2.reshape(Array(1, j, k)).broadcast(Array(i, j, k))).split(1))
    }
    def matmul1Columns(m1: Tensor, m2: Tensor): Seq[Tensor] = {
      val columns1 = m1.split(1)
      m2.split(1).map { column2: Tensor =>
        fold((columns1 zip column2.split(0)).map { case (l, rr) => l * rr.broadcast(l.shape) })
      }
    }
    Seq[(String, () => ByteBuffer)](
      "fill_2x3x5" -> (() => treeBlobOf(Tensor.fill(42.0f, Array(2, 3, 5)))),
      "translate_padding_99" -> (() => treeBlobOf(Tensor.fill(42.0f, Array(2, 3, 5), padding = 99.0f).translate(Array(1, 2, -3)))),
      "translate_of_data" -> (() => treeBlobOf(r(Array(2, 3, 5), 1, 99.0f).translate(Array(1, 2, -3)))),
      "split_last_dimension" -> (() => treeBlobOf(r(Array(1, 1, 1, 2), 2).split(3)(1))),
      "plus_and_multiplication_shared_operand" -> (() => treeBlobOf(doubled * doubled)),
      "c1_tanh_a_times_b_plus_c" -> (() => treeBlobOf(Tensor.tanh(a * b + c))),
      "c2_chain" -> (() => treeBlobOf(Tensor.tanh(Tensor.log(Tensor.exp(a * b + c) + a) * b) + c)),
      "every_operator" -> (() => treeBlobOf(Tensor.min(Tensor.abs(a) / Tensor.sqrt(b), Tensor.max(-a % b, c - a)))),
      "transpose_3d" -> (() => treeBlobOf(r(Array(2, 2, 3), 4).transpose)),
      "c4_permute_translate" -> (() => treeBlobOf(r(Array(4, 5, 6), 7).permute(Array(2, 0, 1)).translate(Array(3, -5, 7)))),
      "broadcast_2x3_to_2x3x4" -> (() => treeBlobOf(r(Array(2, 3), 5).broadcast(Array(2, 3, 4)))),
      "scale_non_integer_coefficients" -> (() => treeBlobOf(r(Array(3, 5), 6).scale(Array(7, 2)))),
      "matmul2_2x3_3x4" -> (() => treeBlobOf(matmul2(r(Array(2, 3), 8), r(Array(3, 4), 9)))),
      "matmul1_join_of_folds" -> (() => treeBlobOf(Tensor.join(matmul1Columns(r(Array(2, 3), 8), r(Array(3, 4), 9))))),
      "axis0_sum_chain" -> (() => treeBlobOf(fold(r(Array(16, 6), 5).split(0)))),
      "axis1_max_chain" -> (() => treeBlobOf(fold(r(Array(6, 16), 5).split(1), Tensor.max(_, _)))),
      "join_fills_at_0" -> (() => treeBlobOf(Tensor.join(Seq(Tensor.fill(42.0f, Array(3, 4)), Tensor.fill(43.0f, Array(3, 4))), 0))),
      "join_fills_at_1" -> (() => treeBlobOf(Tensor.join(Seq(Tensor.fill(42.0f, Array(3, 4)), Tensor.fill(43.0f, Array(3, 4))), 1))),
      "join_fills_at_2" -> (() => treeBlobOf(Tensor.join(Seq(Tensor.fill(42.0f, Array(3, 4)), Tensor.fill(43.0f, Array(3, 4))), 2))),
      "join_of_split_round_trip" -> (() => treeBlobOf(Tensor.join(r(Array(2, 3, 4), 10).split(1)))),
      "non_inline_chain" -> (() => {
        val ni = Tensor.fill(2.0f, Array(2, 3)).nonInline
        treeBlobOf(ni + ni)
      }),
      "sum_of_inline_chain" -> (() => treeBlobOf((a * b + c).reduce(CudaTensors.Monoid.Plus))),
      "min_of_view" -> (() => treeBlobOf(r(Array(4, 5, 6), 7).permute(Array(2, 0, 1)).reduce(CudaTensors.Monoid.Min)))
    )
  }

  for ((name, blob) <- cases) {
    s"$name is written byte for byte like the golden blob" in {
      bytesOf(blob()) should be(golden(name))
    }
  }
}
