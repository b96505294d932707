package com.thoughtworks.compute

import java.nio.{ByteBuffer, ByteOrder}
import java.util.IdentityHashMap

import com.thoughtworks.compute.Trees.AllTrees
import org.lwjgl.system.MemoryUtil

/** Serialises a `Trees` expression graph into the tree blob `cc_compile` / `cc_compile_ex` take
  * (`include/compute_cuda.h`, "expression trees -> kernels"). This is the CUDA backend's counterpart of exporting a tree
  * into `OpenCLKernelBuilder` (`Tensors.scala:1300-1306`): instead of C snippets the "terms" are node indices of a
  * post-order node table, and the code generator lives behind the C ABI.
  *
  * Blob layout (little endian, every field 32 bits unless noted):
  * {{{
  *   u32 magic 'CCT1' = 0x31544343, u32 numberOfNodes, u32 root, u32 outRank, i32 outShape[outRank], then the node records,
  *   children before parents, each `u32 kind` + payload:
  *     1 FloatLiteral    f32 value                                                     Trees.scala:373-380
  *     2 ArrayParameter  u64 id, f32 padding, u32 rank, i32 shape[rank], i32 definitionRoot (-1 = none)   Trees.scala:755-823
  *     3 Transform       u32 array, u32 rows, u32 columns, f64 matrix[rows * columns]  Trees.scala:676-690
  *     4 Extract         u32 array                                                     Trees.scala:660-672
  *     5 Concatenate     u32 n, u32 element[n]                                         Trees.scala:953-973
  *     6 ConcatenateAt   u32 n, u32 position, u32 element[n]       (root only; Tensor.join(tensors, dimension) in one kernel)
  *     10 Exp 11 Log 12 Abs 13 Tanh 14 Sqrt 15 UnaryMinus   u32 operand                Trees.scala:384-470, 560-572
  *     20 Min 21 Max 22 Plus 23 Minus 24 Times 25 Div 26 Percent   u32 lhs, u32 rhs    Trees.scala:455-558
  *     30 Reduce         u32 monoid (22 | 20 | 21 | 24), u32 operand, u32 rank, i32 shape[rank]   (root only, outShape = [])
  * }}}
  *
  * Emission order — normative, `tests/test_scala_twin.py` holds a transliteration of this class to it and compares its
  * output with the blobs of the C++ mirror (`csrc/tensor.cpp`), byte for byte once parameter ids are replaced by their
  * first-visit ordinal:
  *
  *  1. [[write]] walks the tree in post-order, operands left to right, memoised by node IDENTITY exactly like
  *     `Tree.export` memoises in its `ExportContext` (`Trees.scala:201-220`), so a shared sub-DAG is written once. The walk uses
  *     an explicit stack: a per-axis sum over 16384 rows is a 16384-deep chain of `Plus`, far beyond the JVM's default stack
  *     (the reference recurses here, `Trees.scala:70-91, 496-499`).
  *  2. An `ArrayParameter`'s padding is always a `FloatLiteral` (`Tensors.scala:1259`) and is stored inside the record, not as
  *     a node. Its `id` is the producing `Tensor` object (`Tensors.scala:1259`); the blob carries a 64-bit surrogate: 1 + the
  *     ordinal of first emission ([[parameters]] is the side table back to the tensors).
  *  3. Alpha-conversion (`Trees.scala:191-217`) is NOT applied: the library numbers parameters by first visit when it builds its
  *     structural key, which is what alpha-conversion exists for.
  *  4. After the main tree, [[attachDefinition]] may append the closure of a parameter's not-yet-evaluated `InlineTensor` and
  *     point the parameter's `definitionRoot` at it, so that patterns (the split / broadcast / sum matmul,
  *     `benchmarks.scala:188-191`) are matched through the fusion barrier and the i * j * k product is never materialised.
  *
  * @param trees the backend's `trees` instance (`Tensors.scala:224-228`); node classes are path-dependent on it
  */
final class CudaTreeWriter[T <: AllTrees with Singleton](val trees: T) {
  import trees._

  private var body: ByteBuffer = ByteBuffer.allocate(512).order(ByteOrder.LITTLE_ENDIAN)
  private var numberOfNodes = 0
  private val nodeOfTree = new IdentityHashMap[Tree, Integer]

  /** parameter node -> byte offset of its definitionRoot field */
  private val definitionFieldOffset = new java.util.HashMap[Integer, Integer]

  /** The producing tensors (`ArrayParameter.id`) in order of first emission; `id` in the blob = index + 1. */
  val parameters = new java.util.ArrayList[AnyRef]
  private val parameterNode = new IdentityHashMap[AnyRef, Integer]

  private def ensure(numberOfBytes: Int): Unit = {
    if (body.remaining() < numberOfBytes) {
      val larger = ByteBuffer.allocate(math.max(body.capacity() * 2, body.position() + numberOfBytes)).order(ByteOrder.LITTLE_ENDIAN)
      body.flip()
      larger.put(body)
      body = larger
    }
  }

  private def begin(kind: Int, payloadBytes: Int): Int = {
    ensure(4 + payloadBytes)
    body.putInt(kind)
    numberOfNodes += 1
    numberOfNodes - 1
  }

  private def unaryKind(tree: Tree): Int = tree match {
    case _: Exp        => 10
    case _: Log        => 11
    case _: Abs        => 12
    case _: Tanh       => 13
    case _: Sqrt       => 14
    case _: UnaryMinus => 15
    case _             => 0
  }

  private def binaryKind(tree: Tree): Int = tree match {
    case _: Min     => 20
    case _: Max     => 21
    case _: Plus    => 22
    case _: Minus   => 23
    case _: Times   => 24
    case _: Div     => 25
    case _: Percent => 26
    case _          => 0
  }

  /** Operand trees of a node, left to right (`ArrayParameter.padding` is not an operand, see the class comment). */
  private def operands(tree: Tree): Seq[Tree] = tree match {
    case Exp(operand)                  => operand :: Nil
    case Log(operand)                  => operand :: Nil
    case Abs(operand)                  => operand :: Nil
    case Tanh(operand)                 => operand :: Nil
    case Sqrt(operand)                 => operand :: Nil
    case UnaryMinus(operand)           => operand :: Nil
    case UnaryPlus(operand)            => operand :: Nil
    case Min(left, right)              => left :: right :: Nil
    case Max(left, right)              => left :: right :: Nil
    case Plus(left, right)             => left :: right :: Nil
    case Minus(left, right)            => left :: right :: Nil
    case Times(left, right)            => left :: right :: Nil
    case Div(left, right)              => left :: right :: Nil
    case Percent(left, right)          => left :: right :: Nil
    case Extract(array)                => array :: Nil
    case Transform(array, _)           => array :: Nil
    case Concatenate(elementTrees)     => elementTrees
    case _: FloatLiteral               => Nil
    case _: ArrayParameter[_]          => Nil
    case unreachable =>
      // FloatParameter, TupleParameter, Apply, Fill: not constructible through the Tensor API (SURVEY appendix A.1)
      throw new IllegalArgumentException(s"${unreachable.productPrefix} cannot be reached from a Tensor")
  }

  private def emit(tree: Tree): Int = tree match {
    case FloatLiteral(value) =>
      val node = begin(1, 4)
      body.putFloat(value)
      node
    case ArrayParameter(id, FloatLiteral(padding), shape) =>
      val node = begin(2, 8 + 4 + 4 + 4 * shape.length + 4)
      val tensor = id.asInstanceOf[AnyRef]
      parameters.add(tensor)
      parameterNode.put(tensor, node)
      body.putLong(parameters.size.toLong)
      body.putFloat(padding)
      body.putInt(shape.length)
      shape.foreach(body.putInt)
      definitionFieldOffset.put(node, body.position())
      body.putInt(-1)
      node
    case Transform(array, matrix) =>
      val arrayNode: Int = nodeOfTree.get(array)
      val rows = rankOfArray(array)
      if (rows == 0 || matrix.length % rows != 0) {
        throw new IllegalArgumentException(s"a ${matrix.length}-element matrix cannot have $rows rows")
      }
      val node = begin(3, 12 + 8 * matrix.length)
      body.putInt(arrayNode)
      body.putInt(rows)
      body.putInt(matrix.length / rows)
      matrix.foreach(body.putDouble)
      node
    case Extract(array) =>
      val node = begin(4, 4)
      body.putInt(nodeOfTree.get(array))
      node
    case Concatenate(elementTrees) =>
      val node = begin(5, 4 + 4 * elementTrees.length)
      body.putInt(elementTrees.length)
      elementTrees.foreach(element => body.putInt(nodeOfTree.get(element)))
      node
    case UnaryPlus(operand) =>
      nodeOfTree.get(operand) // `+x` is `x` (Tensor.unary_+ returns this, Tensors.scala:900); no record
    case other =>
      val unary = unaryKind(other)
      val binary = binaryKind(other)
      if (unary != 0) {
        val node = begin(unary, 4)
        body.putInt(nodeOfTree.get(operands(other).head))
        node
      } else if (binary != 0) {
        val Seq(left, right) = operands(other)
        val node = begin(binary, 8)
        body.putInt(nodeOfTree.get(left))
        body.putInt(nodeOfTree.get(right))
        node
      } else {
        throw new IllegalArgumentException(s"${other.productPrefix} cannot be reached from a Tensor")
      }
  }

  /** Rows of a Transform's matrix = rank of the array it is applied to. The Tensor API applies `Transform` directly to an
    * `ArrayParameter` (views are pre-composed on the host, `Tensors.scala:979-989`). */
  private def rankOfArray(array: Tree): Int = array match {
    case ArrayParameter(_, _, shape) => shape.length
    case other                       => throw new IllegalArgumentException(s"Transform over ${other.productPrefix}")
  }

  /** Writes `tree` (and whatever of its sub-DAG is not written yet); returns its node index. */
  def write(tree: Tree): Int = {
    val known = nodeOfTree.get(tree)
    if (known != null) {
      known
    } else {
      // (node, expanded?) — a node is emitted when it comes off the stack the second time, after its operands
      val stack = new java.util.ArrayDeque[(Tree, Boolean)]
      stack.push((tree, false))
      while (!stack.isEmpty) {
        val (node, expanded) = stack.pop()
        if (!nodeOfTree.containsKey(node)) {
          val pending = operands(node).filterNot(nodeOfTree.containsKey)
          if (expanded || pending.isEmpty) {
            nodeOfTree.put(node, emit(node))
          } else {
            stack.push((node, true))
            pending.reverseIterator.foreach(operand => stack.push((operand, false)))
          }
        }
      }
      nodeOfTree.get(tree)
    }
  }

  /** kind 6: the root of `Tensor.join(tensors, dimension)` (`Tensors.scala:560-575`) — the element index lands at output
    * dimension `position` instead of last, so the join and the reference's follow-up permute are one kernel. */
  def concatenateAt(elements: Seq[Int], position: Int): Int = {
    val node = begin(6, 8 + 4 * elements.length)
    body.putInt(elements.length)
    body.putInt(position)
    elements.foreach(body.putInt)
    node
  }

  /** kind 30: the root of `Tensor.sum` / `reduce` over an inline operand (`Tensors.scala:673-771` materialises first; this
    * folds the operand's closure inside one kernel). `monoid` = 22 Plus | 20 Min | 21 Max | 24 Times. */
  def reduce(monoid: Int, operand: Int, operandShape: Array[Int]): Int = {
    val node = begin(30, 12 + 4 * operandShape.length)
    body.putInt(monoid)
    body.putInt(operand)
    body.putInt(operandShape.length)
    operandShape.foreach(body.putInt)
    node
  }

  /** The parameter node that stands for `tensor`, if the tree written so far refers to it. */
  def parameterNodeOf(tensor: AnyRef): Option[Int] = Option(parameterNode.get(tensor)).map(_.intValue)

  /** Points the `definitionRoot` of `tensor`'s parameter record at `closure` (written now, after the main tree). */
  def attachDefinition(tensor: AnyRef, closure: Tree): Unit = {
    val node = parameterNode.get(tensor)
    if (node == null) {
      throw new IllegalArgumentException("the tensor is not a parameter of the tree written so far")
    }
    val definitionRoot = write(closure)
    body.putInt(definitionFieldOffset.get(node), definitionRoot)
  }

  /** The finished blob in off-heap memory (the caller frees it with `MemoryUtil.memFree` after `cc_compile_ex`). */
  def finish(root: Int, outShape: Array[Int]): ByteBuffer = {
    val headerBytes = 16 + 4 * outShape.length
    val blob = MemoryUtil.memAlloc(headerBytes + body.position()).order(ByteOrder.LITTLE_ENDIAN)
    blob.putInt(0x31544343)
    blob.putInt(numberOfNodes)
    blob.putInt(root)
    blob.putInt(outShape.length)
    outShape.foreach(blob.putInt)
    val written = body.duplicate()
    written.flip()
    blob.put(written)
    blob.flip()
    blob
  }
}
