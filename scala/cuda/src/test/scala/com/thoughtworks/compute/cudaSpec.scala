package com.thoughtworks.compute

import com.thoughtworks.future._
import com.thoughtworks.raii.asynchronous._
import org.scalatest.{FreeSpec, Matchers}

/** The known answers of the reference's own device specs — `TensorsSpec.scala:27-528`, `cpuSpec.scala:9-38` and the scaladoc
  * examples of `cpu.scala:15-101` — asked of the `cuda` backend. The only change a user makes is the import. (The same answers
  * are held against the C ABI by `tests/test_cuda_goldens.py` in this repository, where no JVM is available.)
  *
  * Needs a B200 and `libcompute_cuda.so` (`-Dcom.thoughtworks.compute.cuda.libname=...`).
  */
class cudaSpec extends FreeSpec with Matchers {
  import cuda._

  private def fold(tensors: Seq[Tensor]): Tensor = tensors.reduce[Tensor](_ + _)

  /** `matrixMultiply` as `TensorsSpec.scala:472-479` / `benchmarks.scala:188-191` write it */
  private def matrixMultiply2(matrix1: Tensor, matrix2: Tensor): Tensor = {
    val Array(i, j) = matrix1.shape
    val Array(`j`, k) = matrix2.shape
    val product = matrix1.broadcast(Array(i, j, k)) * matrix2.reshape(Array(1, j, k)).broadcast(Array(i, j, k))
    fold(product.split(1))
  }

  /** `matrixMultiply` as `TensorsSpec.scala:506-518` / `benchmarks.scala:176-187` write it */
  private def matrixMultiply1(matrix1: Tensor, matrix2: Tensor): Tensor = {
    val columns1 = matrix1.split(1)
    Tensor.join(matrix2.split(1).map { column2: Tensor =>
      fold((columns1 zip column2.split(0)).map { case (l, r) => l * r.broadcast(l.shape) })
    })
  }

  private val m23 = Array(Array(1.0f, 2.0f, 3.0f), Array(4.0f, 5.0f, 6.0f))
  private val m34 = Array(Array(7.0f, 8.0f, 9.0f, 10.0f), Array(11.0f, 12.0f, 13.0f, 14.0f), Array(15.0f, 16.0f, 17.0f, 18.0f))

  "literals print like the reference's (TensorsSpec.scala:57-64, cpu.scala:15-40)" in {
    Tensor(42.0f).toString should be("42.0")
    Tensor(Array(1.0f, 2.0f)).toString should be("[1.0,2.0]")
    Tensor(Array(Seq(1.0f, 2.0f), List(3.0f, 4.0f))).toString should be("[[1.0,2.0],[3.0,4.0]]")
    for (_ <- 0 until 1000) Tensor(42.0f).toString should be("42.0") // TensorsSpec.scala:27-35
  }

  "ragged literals are rejected (TensorsSpec.scala:66-73)" in {
    an[IllegalArgumentException] should be thrownBy Tensor(Seq(Array(1.0f), Array(3.0f, 4.0f)))
  }

  "structurally equal constants share one kernel (TensorsSpec.scala:37-55)" in {
    val filled = Tensor.fill(42.0f, Array(2, 3, 5))
    filled.flatArray.blockingAwait should be(Array.fill(30)(42.0f))
    kernelCache.getIfPresent(filled.getClosure) should not be null
    kernelCache.getIfPresent(Tensor.fill(42.0f, Array(2, 3, 5)).getClosure) should not be null
    kernelCache.getIfPresent(Tensor.fill(43.0f, Array(2, 3, 5)).getClosure) should be(null) // literals are part of the key (Trees.scala:373-380)
  }

  "translate pads what it uncovers (TensorsSpec.scala:75-113)" in {
    val translated = Tensor.fill(42.0f, Array(2, 3, 5), padding = 99.0f).translate(Array(1, 2, -3))
    translated.toString should be(
      "[[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0]]," +
        "[[99.0,99.0,99.0,99.0,99.0],[99.0,99.0,99.0,99.0,99.0],[42.0,42.0,99.0,99.0,99.0]]]")
  }

  "split, plus, times (TensorsSpec.scala:115-138)" in {
    Tensor(Seq(Seq(Seq(Seq(1.0f, 5.0f))))).split(dimension = 3).map(_.toString) should be(Seq("[[[1.0]]]", "[[[5.0]]]"))
    val tensor = Tensor(Seq(Seq(Seq(1.0f, 5.0f))))
    (tensor + tensor).toString should be("[[[2.0,10.0]]]")
    val doubled = tensor + tensor
    (doubled * doubled).toString should be("[[[4.0,100.0]]]")
  }

  "sum (TensorsSpec.scala:251-257)" in {
    Tensor.fill(15625.0f, Array(8, 8)).sum.toString should be("1000000.0")
    Tensor.fill(15625.0f, Array(8, 8)).nonInline.sum.toString should be("1000000.0") // the buffer route (cc_reduce_sum)
  }

  "random is the reference's stream, bit for bit (TensorsSpec.scala:402-409)" in {
    Tensor.random(Array(3, 3), seed = 12345).toString should be(
      "[[0.48931676,0.2949697,0.14271837],[0.9694414,0.26660874,0.07228618],[0.8779875,0.7046564,0.018829918]]")
  }

  "randomNormal is the reference's stream up to the libm (TensorsSpec.scala:259-265, 411-434)" in {
    val expected = Array(1.4561316f, -0.8711971f, -0.7223376f, -2.232667f, -0.24489015f, -0.41490105f, -1.0286478f, -1.392045f, 0.08673929f,
      -0.37037173f, 0.5294154f, -0.5261399f, -0.88834476f, -0.66154f, 0.7035836f, -1.1797824f, -0.93145895f, -1.0812063f, -1.881317f, 0.20438789f,
      -2.5961785f, 1.3082669f, 0.58748704f, -0.01997061f, -1.7090794f, 1.0162057f, 0.33355764f)
    val got = Tensor.randomNormal(Array(3, 3, 3), seed = 54321).flatArray.blockingAwait
    // sqrt / log / cos / sin come from the device's libm: the reference's own CPU drivers disagree in the last printed digit (SURVEY finding 8)
    for ((g, e) <- got zip expected) math.abs(g - e) should be <= 4 * math.ulp(e)
    Tensor.randomNormal(Array.empty[Int], seed = 54321).readScalar.blockingAwait should be(1.4561316f +- 4 * math.ulp(1.4561316f))
  }

  "transpose of ranks 0 to 3 (TensorsSpec.scala:436-466)" in {
    Tensor(42.0f).transpose.toString should be("42.0")
    Tensor(Array(1.0f, 2.0f, 3.0f)).transpose.toString should be("[1.0,2.0,3.0]")
    Tensor(Array(Array(1.0f, 2.0f), Array(3.0f, 4.0f))).transpose.toString should be("[[1.0,3.0],[2.0,4.0]]")
    Tensor(Array(Array(Array(1.0f, 2.0f, 3.0f), Array(4.0f, 5.0f, 6.0f)), Array(Array(7.0f, 8.0f, 9.0f), Array(10.0f, 11.0f, 12.0f)))).transpose.toString should be(
      "[[[1.0,7.0],[4.0,10.0]],[[2.0,8.0],[5.0,11.0]],[[3.0,9.0],[6.0,12.0]]]")
  }

  "broadcast aligns LEADING dimensions (TensorsSpec.scala:491-500)" in {
    Tensor(m23).broadcast(Array(2, 3, 4)).toString should be(
      "[[[1.0,1.0,1.0,1.0],[2.0,2.0,2.0,2.0],[3.0,3.0,3.0,3.0]],[[4.0,4.0,4.0,4.0],[5.0,5.0,5.0,5.0],[6.0,6.0,6.0,6.0]]]")
    Tensor.scalar(42.0f).broadcast(Array(2, 3)).toString should be("[[42.0,42.0,42.0],[42.0,42.0,42.0]]") // cpu.scala:95-100
  }

  "both formulations of matrix multiplication (TensorsSpec.scala:468-489, 502-528)" in {
    val expected = "[[74.0,80.0,86.0,92.0],[173.0,188.0,203.0,218.0]]"
    matrixMultiply2(Tensor(m23), Tensor(m34)).toString should be(expected)
    matrixMultiply1(Tensor(m23), Tensor(m34)).toString should be(expected)
  }

  "the matmul pattern is a tensor-core contraction from 2^25 multiply-adds, and never materialises i * j * k" in {
    val n = 512
    val a = Tensor.random(Array(n, n), seed = 9)
    val b = Tensor.random(Array(n, n), seed = 10)
    val c = matrixMultiply2(a, b).asInstanceOf[InlineTensor]
    c.plan.kind should be(2)
    c.plan.arguments should be(List(a, b)) // the kernel takes A and B; the [n, n, n] product exists only as a definition
  }

  "chained non-inline tensors (cpuSpec.scala:9-16)" in {
    val a = Tensor.fill(2.0f, Array(2, 3)).nonInline
    val b = Tensor.fill(2.0f, Array(2, 3)).nonInline
    val c = (a + b).nonInline
    (c + b).nonInline.toString should be("[[6.0,6.0,6.0],[6.0,6.0,6.0]]")
  }

  "join at every dimension (cpuSpec.scala:18-38, cpu.scala:62-93)" in {
    val a = Tensor.fill(42.0f, Array(3, 4))
    val b = Tensor.fill(43.0f, Array(3, 4))
    val row42 = "[42.0,42.0,42.0,42.0]"
    val row43 = "[43.0,43.0,43.0,43.0]"
    val t0 = Tensor.join(Seq(a, b), 0)
    t0.shape should be(Array(2, 3, 4))
    t0.toString should be(s"[[$row42,$row42,$row42],[$row43,$row43,$row43]]")
    val t1 = Tensor.join(Seq(a, b), 1)
    t1.shape should be(Array(3, 2, 4))
    t1.toString should be(s"[[$row42,$row43],[$row42,$row43],[$row42,$row43]]")
    val t2 = Tensor.join(Seq(a, b), 2)
    t2.shape should be(Array(3, 4, 2))
    t2.toString should be(Seq.fill(3)(Seq.fill(4)("[42.0,43.0]").mkString("[", ",", "]")).mkString("[", ",", "]"))
    val iota = Tensor(Array.tabulate(2, 3, 4)((i, j, k) => (i * 12 + j * 4 + k).toFloat))
    Tensor.join(iota.split(1), 1).toString should be(iota.toString)
  }

  "a per-axis sum of 16384 terms neither overflows the stack nor unrolls (README.md:301-310; the reference recurses, Trees.scala:70-91)" in {
    val x = Tensor.fill(1.0f, Array(16384, 8)).nonInline
    val columnSums = fold(x.split(0)).asInstanceOf[InlineTensor]
    columnSums.plan.kind should be(1)
    columnSums.flatArray.blockingAwait should be(Array.fill(8)(16384.0f))
  }
}
