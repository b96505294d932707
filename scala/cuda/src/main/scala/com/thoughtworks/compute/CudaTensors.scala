package com.thoughtworks.compute

import java.nio.FloatBuffer
import java.util.concurrent.Callable
import java.util.{Collections, IdentityHashMap}

import com.google.common.cache.{AbstractCache, Cache}
import com.thoughtworks.compute.Expressions.{Arrays, Floats, Tuples}
import com.thoughtworks.compute.NDimensionalAffineTransform.MatrixData
import com.thoughtworks.compute.Tensors.{MemoryTrees, TensorBuilder}
import com.thoughtworks.compute.Trees.{AllTrees, StructuralTrees}
import com.thoughtworks.continuation._
import com.thoughtworks.feature.Factory
import com.thoughtworks.future._
import com.thoughtworks.raii.asynchronous._
import com.thoughtworks.raii.covariant._
import com.thoughtworks.tryt.covariant.TryT
import org.lwjgl.system.MemoryUtil
import scalaz.Tags.Parallel
import scalaz.std.list._
import scalaz.syntax.all._
import scalaz.syntax.tag._

import scala.collection.SeqView
import scala.util.{Random, Success, Try}

/** The lazy Tensor API of `Tensors.scala:199-1442` on the B200 backend (`libcompute_cuda.so`).
  *
  * `trait Tensors extends OpenCL` hard-codes both its runtime and its code generator (`Tensors.scala:199, 1301`), so the CUDA
  * backend cannot be mixed into it; this is the sibling trait with the same public surface — `Tensor`, `InlineTensor`,
  * `TransformedTensor`, `NonInlineTensor`, `CachedTensor`, `FillTensor`, every operator, view and slow action — that re-uses
  * the reference's `Trees`, `Expressions`, `NDimensionalAffineTransform`, `Memory` and `Tensors.TensorBuilder` UNCHANGED and
  * replaces what touches the device:
  *
  * | reference (`Tensors.scala`)                       | here                                                                          |
  * |---------------------------------------------------|-------------------------------------------------------------------------------|
  * | `enqueueClosure` (1291-1392): Guava probe, alpha-conversion, `OpenCLKernelBuilder` export, `clBuildProgram`, `clSetKernelArg`, `clEnqueueNDRangeKernel` | [[CudaTreeWriter]] -> `cc_compile_ex` (structural cache + code generator + NVRTC behind the C ABI) -> `cc_launch` |
  * | `kernelCache` (1267-1289)                         | a `Cache` view of the library's structural cache (`cc_kernel_cache_*`)         |
  * | `reduce` + `MonoidPrograms` (303-393, 673-771)    | `cc_reduce_sum` on a buffer; a `Reduce` root when the operand is inline (fused) |
  * | `random` / `randomNormal` programs (398-443)      | `cc_random` / `cc_random_normal` (bit-identical Wang hash / xorshift)          |
  * | `Tensor.apply` (445-463)                          | `cc_buffer_from_host` (asynchronous H2D)                                       |
  * | `toHostBuffer` / `flatBuffer` (285-287, 1099-1109)| `cc_buffer_to_host` into pooled pinned memory; small results are stored into it by the kernel itself |
  *
  * Which half of the C ABI is normative: Scala binds the `cc_*` functions only. The `ct_*` functions are the flat C view of
  * the C++ mirror of this very file (`csrc/tensor.cpp`), which exists because the build environment has no JVM; the three
  * pieces of logic the mirror holds above `cc_*` are re-implemented here, in Scala, and each names its C++ twin:
  * the per-tensor plan ([[CudaTensors#Tensor]]`.plan` = `PlanCache`), definitions across the fusion barrier
  * ([[attachDefinitions]] = `compile_closure`), and direct-to-host small results ([[InlineTensor]]`.flatBuffer` =
  * `Tensor::evaluate_into`).
  */
trait CudaTensors extends Cuda {
  import CudaTensors._

  protected val trees
    : AllTrees with MemoryTrees with StructuralTrees { type Category = Tuples with Floats with Arrays } =
    Factory[AllTrees with MemoryTrees with StructuralTrees].newInstance()

  import trees._

  // ---- buffers in flight (Tensors.scala:253-299) ---------------------------------------------------------------------------

  /** A device buffer and, while its producer may still be running, the event that completes it. */
  protected sealed trait PendingBuffer {
    def buffer: DeviceBuffer
    def eventOption: Option[Event]
    def retain(): Unit
    def release(): Unit

    /** Reads `numberOfFloats` back into pinned host memory that lives as long as the surrounding `Do` scope. */
    def toHostBuffer(numberOfFloats: Int): Do[FloatBuffer] = CudaTensors.this.toHostBuffer(buffer, numberOfFloats, eventOption.toSeq)

    /** the whole buffer (what `PendingBuffer.toHostBuffer` means in the reference, `Tensors.scala:260`) */
    def toHostBuffer: Do[FloatBuffer] = toHostBuffer(buffer.length.toInt)
  }

  protected final case class ReadyBuffer(buffer: DeviceBuffer) extends PendingBuffer {
    def eventOption: Option[Event] = None
    def retain(): Unit = buffer.retain()
    def release(): Unit = buffer.release()
  }

  protected final case class EventBuffer(buffer: DeviceBuffer, event: Event) extends PendingBuffer {
    def eventOption: Option[Event] = Some(event)
    def retain(): Unit = {
      event.retain()
      buffer.retain()
    }
    def release(): Unit = {
      event.release()
      buffer.release()
    }
  }

  // ---- compiled kernels and the kernel cache (Tensors.scala:1263-1289) ------------------------------------------------------

  /** A kernel handle of the library plus the tensors whose buffers it takes, in `cc_launch` order. */
  protected final class CompiledKernel(val handle: Long, val arguments: List[Tensor]) extends MonadicCloseable[UnitContinuation] {
    lazy val info: CudaNative.KernelInfo = CudaNative.kernelInfo(handle)

    /** 0 elementwise, 1 axis reduction, 2 contraction (tcgen05), 3 tiled transpose, 4 whole-tensor fold */
    def kind: Int = info.kind
    def monadicClose: UnitContinuation[Unit] = UnitContinuation.delay { CudaNative.kernelRelease(handle) }
  }

  /** Upper bound on cached kernels, 0 = unbounded like the reference's default `CacheBuilder` (`Tensors.scala:1267-1276`);
    * least-recently-used kernels are evicted beyond it (`cc_kernel_cache_limit`). */
  protected def maximumNumberOfCachedKernels: Long = 0L

  CudaNative.kernelCacheLimit(maximumNumberOfCachedKernels)

  /** The kernel cache, keyed by the STRUCTURE of a term as in the reference (`Trees.scala:23-177, 278-298`): parameters by
    * first-visit ordinal, literals / shapes / paddings / matrices part of the key. The cache itself lives inside the library,
    * which computes the key iteratively from the blob, so probing never recurses to the depth of the tree the way
    * `structuralHashCode` does (a 16384-term per-axis sum overflows a default JVM stack there). This object is the view
    * `TensorsSpec.scala:50-52` looks at. */
  protected[compute] val kernelCache: Cache[ValueTerm, CompiledKernel] = new AbstractCache[ValueTerm, CompiledKernel] {
    def getIfPresent(key: Any): CompiledKernel = key match {
      case term: ValueTerm @unchecked =>
        val writer = new CudaTreeWriter[trees.type](trees)
        val blob = writer.finish(writer.write(term.tree), Array.empty[Int])
        try {
          // the output shape is part of the library's key; a bare term does not know it (the reference's kernels take it at launch,
          // Tensors.scala:1373), so the probe matches any output shape. The handle comes back retained; the probe's caller owns it.
          CudaNative.kernelCacheLookup(blob, anyOutputShape = true) match {
            case 0L     => null
            case handle => new CompiledKernel(handle, Nil)
          }
        } finally MemoryUtil.memFree(blob)
      case _ => null
    }
    override def get(key: ValueTerm, loader: Callable[_ <: CompiledKernel]): CompiledKernel = {
      val cached = getIfPresent(key)
      if (cached != null) cached else loader.call()
    }
    override def size(): Long = CudaNative.kernelCacheSize()
    override def invalidateAll(): Unit = CudaNative.kernelCacheClear()
    override def cleanUp(): Unit = ()
  }

  private def clearCache: UnitContinuation[Unit] = UnitContinuation.execute {
    kernelCache.invalidateAll()
    kernelCache.cleanUp()
  }

  override def monadicClose: UnitContinuation[Unit] = {
    clearCache >> super.monadicClose
  }

  // ---- tree -> kernel --------------------------------------------------------------------------------------------------------

  /** `compile_closure` of the C++ mirror. For every parameter of the main tree whose producing tensor is a not yet evaluated
    * [[InlineTensor]] the tensor's own closure is appended and the parameter's `definitionRoot` points at it (one level deep).
    * The library uses it to look THROUGH the fusion barrier `Tensors.scala:1423-1426` creates — `product.split(1)` makes the
    * inline product a kernel parameter, which evaluated literally is an i * j * k buffer (2 TiB at 8192^3, SURVEY finding 2) —
    * and, when the pattern around it is a reduction or a contraction, composes the definition into the kernel instead.
    * A definition the library has no use for costs nothing: the parameter is then evaluated like any other. */
  private def attachDefinitions(writer: CudaTreeWriter[trees.type]): Unit = {
    val mainParameters = writer.parameters.toArray
    for (id <- mainParameters) {
      id match {
        case inline: InlineTensor @unchecked =>
          writer.attachDefinition(inline, inline.closure.tree)
        case _ =>
      }
    }
  }

  /** Serialises the tree `writeRoot` writes, compiles it (or finds it in the library's structural cache) and resolves the
    * kernel's buffer arguments back to tensors: `cc_compile_ex` reports the blob's parameter ids in ordinal order, and
    * `cc_kernel_arg_param` names the ordinal of each launch argument (a plan may drop, reorder or look through parameters). */
  private def compile(shape: Array[Int])(writeRoot: CudaTreeWriter[trees.type] => Int): CompiledKernel = {
    val writer = new CudaTreeWriter[trees.type](trees)
    val root = writeRoot(writer)
    attachDefinitions(writer)
    val blob = writer.finish(root, shape)
    try {
      val (handle, ids) = CudaNative.compileEx(blob, writer.parameters.size + 1)
      try {
        val numberOfArguments = CudaNative.kernelInfo(handle).numberOfArguments
        val arguments = List.tabulate(numberOfArguments) { i =>
          val ordinal = CudaNative.kernelArgumentParameter(handle, i)
          writer.parameters.get((ids(ordinal) - 1).toInt).asInstanceOf[Tensor]
        }
        new CompiledKernel(handle, arguments)
      } catch {
        case e: Throwable =>
          CudaNative.kernelRelease(handle)
          throw e
      }
    } finally MemoryUtil.memFree(blob)
  }

  /** The blob [[compile]] hands to `cc_compile_ex` for `tensor` (definitions attached), in off-heap memory the caller frees — the hook
    * `CudaTreeWriterSpec` compares with the golden blobs of `tests/golden/tree_blobs/`. */
  private[compute] def treeBlobOf(tensor: Tensor): java.nio.ByteBuffer = {
    val writer = new CudaTreeWriter[trees.type](trees)
    val root = tensor.writeRoot(writer)
    attachDefinitions(writer)
    writer.finish(root, tensor.shape)
  }

  /** enqueueClosure's second half (`Tensors.scala:1331-1381`): evaluate the argument tensors in parallel, allocate the output,
    * launch after the arguments' events. One call, one `cc_launch`; the library picks the stream from the buffers' hazards. */
  private def enqueue(kernel: CompiledKernel, shape: Array[Int], allReduce: Boolean = false): Do[PendingBuffer] = {
    kernel.arguments
      .traverse[ParallelDo, PendingBuffer] { tensor =>
        Parallel(tensor.doBuffer)
      }
      .unwrap
      .flatMap { arguments: List[PendingBuffer] =>
        allocateBuffer(numberOfElements(shape)).flatMap { outputBuffer =>
          Do.monadicCloseable {
              val argumentHandles = arguments.map(_.buffer.handle).toArray
              val waits = arguments.flatMap(_.eventOption.map(_.handle)).toArray
              // a partial sum over the sharded axis is completed by an all-reduce of the freshly written output (cc_shard_launch_allreduce)
              new Event(
                if (allReduce) CudaNative.shardLaunchAllReduce(kernel.handle, argumentHandles, outputBuffer.handle, waits)
                else CudaNative.launch(kernel.handle, argumentHandles, outputBuffer.handle, waits))
            }
            .map { event =>
              EventBuffer(outputBuffer, event): PendingBuffer
            }
        }
      }
  }

  // ---- gathers of sharded tensors ------------------------------------------------------------------------------------------------------

  /** What is cached per communicator (`GatherTensor::PerCommunicator` in tensor.cpp): the block sizes all ranks agreed on, and one
    * symmetric (peer-mapped) arena per size — `cuMemAlloc` + an IPC handle exchange, far too slow per evaluation; the entry barrier of
    * `cc_shard_launch_allgather` protects an arena against the previous gather's readers. Collective evaluations run one at a time. */
  private object communicatorCache {
    private var generation = 0L
    private val agreed = scala.collection.mutable.Set.empty[Long]
    private val arenas = scala.collection.mutable.Map.empty[Long, DeviceBuffer]
    private def current(): Unit = {
      val now = CudaNative.commGeneration()
      if (now != generation) { // a new communicator: the old arenas' memory went with the old one
        arenas.values.foreach(_.release())
        arenas.clear()
        agreed.clear()
        generation = now
      }
    }
    def requireEqualBlocks(blockFloats: Long): Unit = synchronized {
      current()
      if (!agreed(blockFloats)) {
        if (!CudaNative.shardAgree(blockFloats)) {
          throw new IllegalArgumentException(
            s"gather needs equal row blocks on every rank (this rank holds $blockFloats floats): pad the leading axis to a multiple of the number of ranks")
        }
        agreed += blockFloats
      }
    }
    def arena(numberOfFloats: Long): DeviceBuffer = synchronized {
      current()
      arenas.getOrElseUpdate(numberOfFloats, new DeviceBuffer(CudaNative.commSymmetricAlloc(numberOfFloats)))
    }
  }

  /** A contraction over a row block, gathered from its own epilogue (`cc_shard_launch_allgather` -> `cc_matmul_3xtf32_allgather`):
    * every accumulator tile is TMA-stored into every rank's copy of the symmetric arena, so the exchange overlaps the MMAs. */
  private def gatherFromEpilogue(block: InlineTensor, wholeFloats: Long, zeroCopy: Boolean): Do[PendingBuffer] = {
    val arena = communicatorCache.arena(wholeFloats)
    block.plan.arguments
      .traverse[ParallelDo, PendingBuffer](tensor => Parallel(tensor.doBuffer))
      .unwrap
      .flatMap { arguments: List[PendingBuffer] =>
        Do.monadicCloseable {
            val (event, _) = CudaNative.shardLaunchAllGather(block.plan.handle,
                                                             arguments.map(_.buffer.handle).toArray,
                                                             arena.handle,
                                                             arguments.flatMap(_.eventOption.map(_.handle)).toArray)
            new Event(event)
          }
          .flatMap { gathered =>
            if (zeroCopy) {
              arena.retain()
              Do.monadicCloseable(new DeviceBuffer(arena.handle)).map(view => EventBuffer(view, gathered): PendingBuffer)
            } else {
              allocateBuffer(wholeFloats).flatMap { copy =>
                Do.monadicCloseable(new Event(CudaNative.bufferCopy(copy.handle, arena.handle, wholeFloats, Array(gathered.handle))))
                  .map(copied => EventBuffer(copy, copied): PendingBuffer)
              }
            }
          }
      }
  }

  // ---- the Tensor companion (Tensors.scala:395-600) ----------------------------------------------------------------------------

  object Tensor {

    /** A tensor of the given (nested) elements, copied to the device on first use (`Tensors.scala:445-463`). */
    def apply[A](elements: A, padding: Float = 0.0f)(implicit tensorBuilder: TensorBuilder.Aux[A, Float]): NonInlineTensor = {
      val shape0 = tensorBuilder.shape(elements).toArray
      val padding0 = padding
      new NonInlineTensor {
        val shape: Array[Int] = shape0
        val padding: Float = padding0
        private[compute] lazy val doBuffer: Do[PendingBuffer] = {
          Do(TryT(ResourceT(UnitContinuation.delay {
            val data = tensorBuilder.flatten(elements).toArray
            val hostBuffer = MemoryUtil.memAllocFloat(data.length)
            hostBuffer.duplicate().put(data)
            Resource(value = Success(hostBuffer): Try[FloatBuffer], release = UnitContinuation.delay { MemoryUtil.memFree(hostBuffer) })
          }))).flatMap { hostBuffer =>
            allocateBufferFrom(hostBuffer).flatMap {
              case (deviceBuffer, copied) =>
                // the host staging memory is freed when this scope closes: wait for the copy first
                Do.garbageCollected(waitForComplete(copied)).map { _: Unit =>
                  ReadyBuffer(deviceBuffer): PendingBuffer
                }
            }
          }.shared
        }
      }
    }

    def scalar(value: Float, padding: Float = 0.0f): InlineTensor = fill(value, ScalarShape, padding)

    /** A constant tensor: a literal inlined into whichever kernel uses it, and part of that kernel's cache key
      * (`Tensors.scala:469-477, 1394-1397`; `Trees.scala:373-380`). */
    def fill(value: Float, shape: Array[Int], padding: Float = 0.0f): InlineTensor = {
      val (value0, shape0, padding0) = (value, shape, padding)
      new FillTensor {
        val value: Float = value0
        val shape: Array[Int] = shape0
        val padding: Float = padding0
      }
    }

    private def generated(shape0: Array[Int], padding0: Float)(generate: (Long, Long) => Long): NonInlineTensor = {
      new NonInlineTensor {
        val shape: Array[Int] = shape0
        val padding: Float = padding0
        private[compute] lazy val doBuffer: Do[PendingBuffer] = {
          val size = numberOfElements(shape)
          allocateBuffer(size + (size & 1L)).flatMap { buffer =>
            Do.monadicCloseable(new Event(generate(buffer.handle, size))).map { event =>
              EventBuffer(buffer, event): PendingBuffer
            }
          }.shared
        }
      }
    }

    /** Uniform [0, 1): `buffer[i] = wang_hash(i ^ seed) / 2^32`, bit-identical to the reference's kernel
      * (`Tensors.scala:106-117, 432-443, 479-497`; golden `TensorsSpec.scala:405-406`). */
    def random(shape: Array[Int], seed: Int = Random.nextInt(), padding: Float = 0.0f): NonInlineTensor =
      generated(shape, padding)((buffer, size) => CudaNative.random(buffer, size, seed))

    /** Box-Muller pairs from `hash(i ^ seed)` and `xorshift` of it (`Tensors.scala:398-429, 500-524`). */
    def randomNormal(shape: Array[Int], seed: Int = Random.nextInt(), padding: Float = 0.0f): NonInlineTensor =
      generated(shape, padding)((buffer, size) => CudaNative.randomNormal(buffer, size, seed))

    def abs(operand: Tensor): InlineTensor = operand.unary(trees.float.abs(_))
    def sqrt(operand: Tensor): InlineTensor = operand.unary(trees.float.sqrt(_))
    def tanh(operand: Tensor): InlineTensor = operand.unary(trees.float.tanh(_))
    def exp(operand: Tensor): InlineTensor = operand.unary(trees.float.exp(_))
    def log(operand: Tensor): InlineTensor = operand.unary(trees.float.log(_))

    def min(leftHandSide: Tensor, rightHandSide: Tensor): InlineTensor =
      leftHandSide.binary(rightHandSide)(trees.float.min(_, _))

    def max(leftHandSide: Tensor, rightHandSide: Tensor): InlineTensor =
      leftHandSide.binary(rightHandSide)(trees.float.max(_, _))

    private def forced[A](seq: Seq[A]): Seq[A] = seq match {
      case view: SeqView[A, _] @unchecked => view.force[A, Seq[A]](collection.breakOut)
      case strict                         => strict
    }

    private def joined(tensors: Seq[Tensor], position: Option[Int]): NonInlineTensor = {
      val head = tensors.head
      val rank = head.shape.length
      val joinedShape = position match {
        case Some(dimension) => head.shape.patch(dimension, Array(tensors.length), 0)
        case None            => head.shape :+ tensors.length
      }
      val anyRowBlock = tensors.exists(_.distribution == Distribution.RowBlock)
      if (anyRowBlock && position.contains(0)) {
        throw new UnsupportedOperationException("join at dimension 0 would put the new dimension in front of the sharded leading axis: gather first")
      }
      new NonInlineTensor {
        val shape: Array[Int] = joinedShape
        val padding: Float = head.padding
        // the new dimension is never the leading one here: row blocks stay row blocks
        override val distribution: Distribution = if (anyRowBlock) Distribution.RowBlock else Distribution.Whole
        private[compute] override def writeRoot(writer: CudaTreeWriter[trees.type]): Int = {
          val elements = tensors.map(tensor => writer.write(tensor.closure.tree))
          position match {
            case Some(dimension) if dimension != rank => writer.concatenateAt(elements, dimension)
            case _                                    => writer.write(trees.tuple.join(tensors.map(_.closure): _*).tree)
          }
        }
        private[compute] lazy val plan: CompiledKernel = compile(joinedShape)(writeRoot)
        private[compute] lazy val doBuffer: Do[PendingBuffer] = Do.suspend(enqueue(plan, joinedShape)).shared
      }
    }

    /** `tensors` become the slices of a new LAST dimension (`Tensors.scala:577-598`): one kernel whose root is
      * `Concatenate(elements)` (`Trees.scala:953-973`). */
    def join(tensors0: Seq[Tensor]): NonInlineTensor = joined(forced(tensors0).map(_.combined), None)

    /** `tensors` become the slices of a new dimension at `dimension`. The reference joins last and then gathers a permuted
      * view of the result in a second kernel (`Tensors.scala:560-575`); here the element index is placed at `dimension` by the
      * same kernel (`ConcatenateAt`), with the same values and shape. */
    def join(tensors0: Seq[Tensor], dimension: Int): Tensor = {
      val tensors = forced(tensors0).map(_.combined)
      if (dimension < 0 || dimension > tensors.head.shape.length) {
        throw new IllegalArgumentException
      }
      joined(tensors, Some(dimension))
    }
  }

  trait CachedTensor extends NonInlineTensor

  // ---- Tensor (Tensors.scala:636-1261) ---------------------------------------------------------------------------------------------

  sealed trait Tensor { thisTensor =>

    /** @group metadata */
    def shape: Array[Int]

    /** @group metadata */
    def padding: Float

    /** How this tensor is spread over the ranks of the communicator; see [[shard]]. Follows the tensor through the lazy graph.
      * @group metadata */
    def distribution: Distribution = Distribution.Whole

    protected[compute] val closure: FloatTerm

    private[CudaTensors] def floatClosure: FloatTerm = closure

    /** The back door to [[closure]] for tests (`Tensors.scala:1090`). */
    private[compute] def getClosure: FloatTerm = closure

    private[compute] def doBuffer: Do[PendingBuffer]

    /** Writes the tree this tensor's own kernel is compiled from and returns its root: the closure for most tensors, a `Concatenate` /
      * `ConcatenateAt` root for joins, a `Reduce` root for folds of inline operands. */
    private[compute] def writeRoot(writer: CudaTreeWriter[trees.type]): Int = writer.write(closure.tree)

    /** `array.parameter(this, float.literal(padding), shape)` (`Tensors.scala:1253-1260`): this tensor as a kernel parameter. */
    @transient
    private[CudaTensors] lazy val arrayTerm = {
      if (shape == null) {
        throw new IllegalArgumentException
      }
      array.parameter(this, float.literal(padding), shape)
    }

    /** Pins the evaluated buffer until the returned resource is released (`Tensors.scala:642-666`). */
    def doCache: Do[CachedTensor] = {
      doBuffer.intransitiveFlatMap { pendingBuffer =>
        Do.resource {
          pendingBuffer.retain()
          val cached: CachedTensor = new CachedTensor {
            val shape: Array[Int] = thisTensor.shape
            val padding: Float = thisTensor.padding
            // (a partial sum comes back from doBuffer all-reduced: whole)
            override val distribution: Distribution =
              if (thisTensor.distribution == Distribution.RowBlock) Distribution.RowBlock else Distribution.Whole
            private[compute] val doBuffer: Do[PendingBuffer] = Do.resource {
              pendingBuffer.retain()
              Resource(pendingBuffer, UnitContinuation.delay { pendingBuffer.release() })
            }
          }
          Resource(cached, UnitContinuation.delay { pendingBuffer.release() })
        }
      }
    }

    /** @group delayed */
    def nonInline: NonInlineTensor

    // ---- tensors sharded over the GPUs of one box (include/compute_cuda.h: ct_shard / ct_gather; C++ twin: tensor.cpp) --------------

    /** Declares this tensor (`[rows on this rank, ...]`) to be THIS rank's row block of a tensor sharded along its leading axis.
      * User code stays the single-GPU code: elementwise operators, views that keep the leading axis in place and operations with
      * replicated operands keep a row block a row block (no exchange; the split / broadcast / sum matmul of a row block of A with a
      * replicated B is still one tcgen05 contraction per rank); `sum` is the GLOBAL sum; `split(0)` yields the local rows as partial
      * contributions whose fold is all-reduced when it is evaluated or used by anything but `+`; views that would mix the sharded
      * axis throw (gather or replicate first, SURVEY §8e). Collective operations must run in the same order on every rank.
      * @group delayed */
    def shard: NonInlineTensor = {
      if (distribution != Distribution.Whole || shape.isEmpty) {
        throw new IllegalArgumentException("only a whole, non-scalar tensor can be declared a row block")
      }
      new NonInlineTensor {
        val shape: Array[Int] = thisTensor.shape
        val padding: Float = thisTensor.padding
        override def distribution: Distribution = Distribution.RowBlock
        private[compute] def doBuffer: Do[PendingBuffer] = thisTensor.doBuffer
      }
    }

    /** partial sum -> the sum itself (a fusion barrier: evaluating a partial sum all-reduces it); identity for anything else */
    private[CudaTensors] def combined: Tensor = {
      if (distribution != Distribution.PartialSum) this
      else
        new NonInlineTensor {
          val shape: Array[Int] = thisTensor.shape
          val padding: Float = thisTensor.padding
          private[compute] def doBuffer: Do[PendingBuffer] = thisTensor.doBuffer
        }
    }

    /** The whole tensor on every rank (`[numberOfRanks * rows, ...]`; needs equal blocks — checked collectively, uneven blocks are an
      * `IllegalArgumentException` on every rank). A sharded matmul result is gathered by the contraction's own epilogue (TMA stores
      * into every rank's copy over NVLink). `zeroCopy` returns a view of the communicator's symmetric arena, valid until the next
      * gather of the same size; otherwise the result is copied out of it.
      * @group delayed */
    def gather(zeroCopy: Boolean = false): Tensor = distribution match {
      case Distribution.Whole      => this
      case Distribution.PartialSum => combined
      case Distribution.RowBlock =>
        val (numberOfRanks, _) = CudaNative.commInfo()
        val blockFloats = numberOfElements(shape)
        new NonInlineTensor {
          val shape: Array[Int] = (thisTensor.shape.head * numberOfRanks) +: thisTensor.shape.tail
          val padding: Float = thisTensor.padding
          private[compute] lazy val doBuffer: Do[PendingBuffer] = Do.suspend {
            if (numberOfRanks == 1) {
              thisTensor.doBuffer
            } else {
              communicatorCache.requireEqualBlocks(blockFloats)
              thisTensor match {
                case inline: InlineTensor
                    if CudaNative.commPeerEnabled() && inline.plan.kind == 2 && inline.shape.length == 2 && inline.shape(1) % 4 == 0 =>
                  gatherFromEpilogue(inline, blockFloats * numberOfRanks, zeroCopy)
                case inline: InlineTensor if CudaNative.commPeerEnabled() =>
                  // any other kernel of the block's own (twin: GatherTensor::evaluate in tensor.cpp): the library launches it and gathers —
                  // in ONE launch when the kernel can collect the other ranks' values itself (a row-owner reduction:
                  // `shard.split(1).reduce(_ + _).gather`), else kernel + one-shot all-gather
                  inline.plan.arguments
                    .traverse[ParallelDo, PendingBuffer](tensor => Parallel(tensor.doBuffer))
                    .unwrap
                    .flatMap { arguments: List[PendingBuffer] =>
                      allocateBuffer(blockFloats * numberOfRanks).flatMap { whole =>
                        Do.monadicCloseable {
                            val (event, _) = CudaNative.shardLaunchAllGather(inline.plan.handle,
                                                                             arguments.map(_.buffer.handle).toArray,
                                                                             whole.handle,
                                                                             arguments.flatMap(_.eventOption.map(_.handle)).toArray)
                            new Event(event)
                          }
                          .map(event => EventBuffer(whole, event): PendingBuffer)
                      }
                    }
                case _ =>
                  thisTensor.doBuffer.flatMap { block =>
                    allocateBuffer(blockFloats * numberOfRanks).flatMap { whole =>
                      Do.monadicCloseable(new Event(CudaNative.allGather(block.buffer.handle, whole.handle, blockFloats, block.eventOption.map(_.handle).toArray)))
                        .map(event => EventBuffer(whole, event): PendingBuffer)
                    }
                  }
              }
            }
          }.shared
        }
    }

    /** `Tensor.sum` and its siblings (`Tensors.scala:303-393, 673-771`; `MonoidPrograms` is generic over append / zero).
      * A materialised operand summed with `+` runs the library's reduction program over the buffer (`cc_reduce_sum`); an
      * inline operand is folded INSIDE one kernel (a `Reduce` root: one pass over the inputs, nothing materialised — the
      * reference always materialises first, `Tensors.scala:678`). Both fold in the same order, so `e.sum` and
      * `e.doCache.sum` agree bit for bit. */
    def reduce(monoid: Monoid): NonInlineTensor = {
      if (distribution == Distribution.PartialSum) {
        return combined.reduce(monoid)
      }
      // a row block folds to the GLOBAL result: local fold + all-reduce of its one float (only + is combined across ranks)
      val acrossRanks = distribution == Distribution.RowBlock
      if (acrossRanks && monoid != Monoid.Plus) {
        throw new UnsupportedOperationException("only + is combined across ranks: gather a row block before reducing it with another monoid")
      }
      new NonInlineTensor {
        val shape: Array[Int] = ScalarShape
        val padding: Float = thisTensor.padding
        private[compute] override def writeRoot(writer: CudaTreeWriter[trees.type]): Int =
          writer.reduce(monoid.kind, writer.write(thisTensor.closure.tree), thisTensor.shape)
        private[compute] lazy val doBuffer: Do[PendingBuffer] = {
          thisTensor match {
            case buffered: NonInlineTensor if monoid == Monoid.Plus =>
              buffered.doBuffer.flatMap { input =>
                allocateBuffer(1L).flatMap { output =>
                  Do.monadicCloseable {
                      val waits = input.eventOption.map(_.handle).toArray
                      val length = numberOfElements(thisTensor.shape)
                      // ONE kernel folds the block and all-reduces the result over NVLink peer memory (cc_reduce_sum_allreduce)
                      new Event(
                        if (acrossRanks) CudaNative.reduceSumAllReduce(input.buffer.handle, length, output.handle, waits)
                        else CudaNative.reduceSum(input.buffer.handle, length, output.handle, waits))
                    }
                    .map { event =>
                      EventBuffer(output, event): PendingBuffer
                    }
                }
              }
            case _ =>
              Do.suspend {
                enqueue(compile(ScalarShape)(writeRoot), ScalarShape, allReduce = acrossRanks)
              }
          }
        }.shared
      }
    }

    /** @group delayed */
    def sum: NonInlineTensor = reduce(Monoid.Plus)

    /** @group slow */
    override def toString: String = {
      flatArray.map { values =>
        def render(dimensions: List[Int], from: Int, until: Int): String = dimensions match {
          case Nil =>
            if (until - from != 1) {
              throw new IllegalArgumentException(s"shape${shape.mkString("(", ",", ")")} does not match the data size (${values.length})")
            }
            values(from).toString
          case size :: rest =>
            val step = if (size == 0) 0 else (until - from) / size
            (0 until size).map(i => render(rest, from + i * step, from + (i + 1) * step)).mkString("[", ",", "]")
        }
        render(shape.toList, 0, values.length)
      }.blockingAwait
    }

    // ---- shapes of binary operations (Tensors.scala:203-222): LEADING dimensions align, missing / unit ones stretch ----

    private[CudaTensors] def binary(rightHandSide: Tensor, keepsPartialSums: Boolean = false)(
        operator: (FloatTerm, FloatTerm) => FloatTerm): InlineTensor = {
      // partial sums stay partial only under + with another partial sum (the fold of split(0) slices); anything else needs the sum itself
      val partial = keepsPartialSums && distribution == Distribution.PartialSum && rightHandSide.distribution == Distribution.PartialSum
      val (leftOperand, rightOperand) = if (partial) (this, rightHandSide) else (combined, rightHandSide.combined)
      val commonShape = autoBroadcastShape(leftOperand.shape, rightOperand.shape)
      val left = leftOperand.broadcast(commonShape)
      val right = rightOperand.broadcast(commonShape)
      // a row block combined with a replicated operand (B of the row-sharded matmul, a constant) is a row block
      val result =
        if (partial) Distribution.PartialSum
        else if (left.distribution == Distribution.RowBlock || right.distribution == Distribution.RowBlock) Distribution.RowBlock
        else Distribution.Whole
      left.derivedTensor(operator(left.floatClosure, right.floatClosure), result)
    }

    private[CudaTensors] def unary(operator: FloatTerm => FloatTerm): InlineTensor = {
      val operand = combined // f(partial sum) needs the sum
      operand.derivedTensor(operator(operand.floatClosure), operand.distribution)
    }

    private[CudaTensors] def derivedTensor(newClosure: FloatTerm, newDistribution: Distribution): InlineTensor = {
      new InlineTensor {
        val shape: Array[Int] = thisTensor.shape
        val padding: Float = thisTensor.padding
        override val distribution: Distribution = newDistribution
        protected[compute] val closure: FloatTerm = newClosure
      }
    }

    /** @group delayed */
    def unary_- : InlineTensor = unary(-_)

    /** @group delayed */
    def unary_+ : this.type = this

    /** @group delayed */
    def +(rightHandSide: Tensor): InlineTensor = binary(rightHandSide, keepsPartialSums = true)(_ + _)

    /** @group delayed */
    def -(rightHandSide: Tensor): InlineTensor = binary(rightHandSide)(_ - _)

    /** @group delayed */
    def *(rightHandSide: Tensor): InlineTensor = binary(rightHandSide)(_ * _)

    /** @group delayed */
    def /(rightHandSide: Tensor): InlineTensor = binary(rightHandSide)(_ / _)

    /** Scala's `Float %` (the backend emits `fmodf`; the reference's OpenCL C emits `%` on floats, which is not valid
      * OpenCL C, `OpenCLKernelBuilder.scala:564-570`).
      * @group delayed */
    def %(rightHandSide: Tensor): InlineTensor = binary(rightHandSide)(_ % _)

    // ---- views (Tensors.scala:816-1074): row-major matrices, one row per dimension of the data viewed, one column per
    // dimension of the view plus a constant column; chains of views are pre-multiplied so any chain costs ONE gather ----

    /** A view whose index `g` reads `this[matrix * (g, 1)]`, or `padding` outside (`Tensors.scala:978-1003`). */
    private[CudaTensors] def transform(newShape: Array[Int], viewToThis: MatrixData): TransformedTensor = {
      if (distribution == Distribution.PartialSum) {
        return combined.transform(newShape, viewToThis) // (the padding of a view would be added once per rank)
      }
      val viewDistribution = if (distribution == Distribution.RowBlock) viewOfRowBlock(shape, newShape, viewToThis) else Distribution.Whole
      thisTensor match {
        case view: TransformedTensor =>
          val composed = NDimensionalAffineTransform.preConcatenate(viewToThis, view.matrix, newShape.length)
          new TransformedTensor {
            val checkpoint: Tensor = view.checkpoint
            val matrix: MatrixData = composed
            val shape: Array[Int] = newShape
            val padding: Float = view.padding
            override val distribution: Distribution = viewDistribution
          }
        case _ =>
          new TransformedTensor {
            val checkpoint: Tensor = thisTensor
            val matrix: MatrixData = viewToThis
            val shape: Array[Int] = newShape
            def padding: Float = checkpoint.padding
            override val distribution: Distribution = viewDistribution
          }
      }
    }

    /** @group delayed */
    def broadcast(newShape: Array[Int]): Tensor = {
      if (java.util.Arrays.equals(newShape, shape)) {
        this
      } else {
        thisTensor match {
          case constant: FillTensor =>
            Tensor.fill(constant.value, newShape, constant.padding) // a broadcast constant stays a constant (Tensors.scala:820-826)
          case _ =>
            val rank = shape.length
            val columns = newShape.length + 1
            val matrix = new Array[Double](rank * columns)
            for (i <- 0 until rank) {
              if (i < newShape.length && shape(i) == newShape(i)) {
                matrix(i * columns + i) = 1.0
              } else if (shape(i) != 1) {
                throw new IllegalArgumentException(s"Cannot broadcast ${shape.mkString("[", ",", "]")} to ${newShape.mkString("[", ",", "]")}")
              }
            }
            transform(newShape, matrix)
        }
      }
    }

    /** Reinterprets the row-major data with a new shape; an inline receiver is materialised (`Tensors.scala:879-888`).
      * @group delayed */
    def reshape(newShape: Array[Int]): NonInlineTensor = {
      if (numberOfElements(newShape) != numberOfElements(shape)) {
        throw new IllegalArgumentException
      }
      if (distribution == Distribution.PartialSum) {
        return combined.reshape(newShape)
      }
      if (distribution == Distribution.RowBlock && (newShape.isEmpty || newShape(0) != shape(0))) {
        throw new UnsupportedOperationException("reshape of a row block changes the sharded leading axis: gather first")
      }
      new NonInlineTensor {
        val shape: Array[Int] = newShape
        val padding: Float = thisTensor.padding
        override def distribution: Distribution = thisTensor.distribution
        private[compute] def doBuffer: Do[PendingBuffer] = thisTensor.doBuffer
      }
    }

    /** Nearest-neighbour resampling to `newShape`: coefficient `shape(i) / newShape(i)` on the diagonal (`Tensors.scala:950-965`);
      * non-integer coefficients take the reference's exact route in the backend (3 fraction digits, double arithmetic, `(int)`
      * truncation: `OpenCLKernelBuilder.scala:14-32, 386`).
      * @group delayed */
    def scale(newShape: Array[Int]): TransformedTensor = {
      val rank = newShape.length
      if (rank != shape.length) {
        throw new IllegalArgumentException
      }
      val matrix = new Array[Double](rank * (rank + 1))
      for (i <- 0 until rank) {
        matrix(i * (rank + 1) + i) = shape(i).toDouble / newShape(i)
      }
      transform(newShape, matrix)
    }

    /** `out[g] = this[g - offset]`, `padding` where that leaves the data (`Tensors.scala:970-976`).
      * @group delayed */
    def translate(offset: Array[Double], newShape: Array[Int] = shape): TransformedTensor = {
      if (offset.length != shape.length) {
        throw new IllegalArgumentException
      }
      transform(newShape, NDimensionalAffineTransform.translate(offset.map(-_)))
    }

    /** Dimension `n` of the result is dimension `dimensions(n)` of this tensor (`Tensors.scala:1008-1025`).
      * @group delayed */
    def permute(dimensions: Array[Int]): TransformedTensor = {
      val rank = shape.length
      if (dimensions.length != rank) {
        throw new IllegalArgumentException
      }
      val matrix = new Array[Double](rank * (rank + 1))
      for ((oldDimension, newDimension) <- dimensions.zipWithIndex) {
        matrix(oldDimension * (rank + 1) + newDimension) = 1.0
      }
      transform(dimensions.map(shape(_)), matrix)
    }

    /** @group delayed */
    def transpose: TransformedTensor = permute(shape.indices.reverse.toArray)

    /** The `shape(dimension)` slices along `dimension`, each a view with that dimension removed (`Tensors.scala:1035-1074`).
      * Folding them — `t.split(axis).reduce(_ + _)` (`README.md:301-310`), the matmul formulations of
      * `benchmarks.scala:174-193` — builds an unrolled chain the backend re-rolls into a real reduction / contraction.
      * @group delayed */
    def split(dimension: Int): IndexedSeq[TransformedTensor] = {
      val rank = shape.length
      val sliceShape = shape.patch(dimension, Nil, 1)
      new IndexedSeq[TransformedTensor] {
        override def stringPrefix = "TensorSeq"
        val length: Int = shape(dimension)
        def apply(index: Int): TransformedTensor = {
          // rows: dimensions of this tensor; columns: the rank - 1 dimensions of the slice + the constant
          val matrix = new Array[Double](rank * rank)
          for (row <- 0 until rank) {
            if (row < dimension) matrix(row * rank + row) = 1.0
            else if (row == dimension) matrix(row * rank + rank - 1) = index.toDouble
            else matrix(row * rank + row - 1) = 1.0
          }
          transform(sliceShape, matrix)
        }
      }
    }

    // ---- slow actions (Tensors.scala:1099-1246) ---------------------------------------------------------------------------------

    /** Evaluates and reads back into off-heap memory, row-major. The memory is pinned and pooled by the library and valid
      * only inside the `Do` scope, like the LWJGL buffer of the reference (`OpenCL.scala:691-696`).
      * @group slow */
    def flatBuffer: Do[FloatBuffer] = {
      doBuffer.intransitiveFlatMap { pendingBuffer =>
        pendingBuffer.toHostBuffer(numberOfElements(shape).toInt)
      }
    }

    /** @group slow */
    def flatArray: Future[Array[Float]] = {
      flatBuffer.intransitiveMap(Memory.FloatMemory.toArray).run
    }

    /** @group slow */
    def readScalar: Future[Float] = flatArray.map(_(0))

    /** @group slow */
    def read1DArray: Future[Array[Float]] = flatArray

    /** @group slow */
    def read2DArray: Future[Array[Array[Float]]] = flatArray.map(nested2(_).map(_.toArray).toArray)

    /** @group slow */
    def read3DArray: Future[Array[Array[Array[Float]]]] = flatArray.map(nested3(_).map(_.map(_.toArray).toArray).toArray)

    /** @group slow */
    def read4DArray: Future[Array[Array[Array[Array[Float]]]]] =
      flatArray.map(nested4(_).map(_.map(_.map(_.toArray).toArray).toArray).toArray)

    /** @group slow */
    def read5DArray: Future[Array[Array[Array[Array[Array[Float]]]]]] =
      flatArray.map(nested5(_).map(_.map(_.map(_.map(_.toArray).toArray).toArray).toArray).toArray)

    /** @group slow */
    def read1DSeq: Future[Seq[Float]] = flatArray.map(_.toSeq)

    /** @group slow */
    def read2DSeq: Future[Seq[Seq[Float]]] = flatArray.map(nested2)

    /** @group slow */
    def read3DSeq: Future[Seq[Seq[Seq[Float]]]] = flatArray.map(nested3)

    /** @group slow */
    def read4DSeq: Future[Seq[Seq[Seq[Seq[Float]]]]] = flatArray.map(nested4)

    /** @group slow */
    def read5DSeq: Future[Seq[Seq[Seq[Seq[Seq[Float]]]]]] = flatArray.map(nested5)

    // row-major regrouping of the flat array, innermost dimension last (Tensors.scala:1121-1155)
    private def grouped[A](flat: Seq[A], innerSize: Int): Seq[Seq[A]] =
      if (innerSize == 0) Seq.fill(shape(0))(Nil) else flat.grouped(innerSize).toVector
    private def nested2(flat: Array[Float]): Seq[Seq[Float]] = grouped(flat.toSeq, shape.last)
    private def nested3(flat: Array[Float]): Seq[Seq[Seq[Float]]] = grouped(nested2(flat), shape(shape.length - 2))
    private def nested4(flat: Array[Float]): Seq[Seq[Seq[Seq[Float]]]] = grouped(nested3(flat), shape(shape.length - 3))
    private def nested5(flat: Array[Float]): Seq[Seq[Seq[Seq[Seq[Float]]]]] = grouped(nested4(flat), shape(shape.length - 4))
  }

  // ---- the three kinds of tensor (Tensors.scala:1394-1440) ----------------------------------------------------------------------------

  /** An intermediate expression that is merged into whichever kernel uses it (`Tensors.scala:1399-1411`). */
  trait InlineTensor extends Tensor { thisInlineTensor =>

    /** `PlanCache` of the C++ mirror: the graph under a tensor is immutable, so the kernel its closure compiles to and the
      * tensors that kernel takes never change; they are resolved once per tensor, after which a slow action costs one
      * `cc_launch` and no tree walk (the reference re-hashes the whole tree on every evaluation, `Tensors.scala:1293`). */
    @transient
    private[compute] lazy val plan: CompiledKernel = compile(shape)(writeRoot)

    private[compute] lazy val doBuffer: Do[PendingBuffer] =
      Do.suspend(enqueue(plan, shape, allReduce = distribution == Distribution.PartialSum)).shared

    def nonInline: NonInlineTensor = new NonInlineTensor {
      val shape: Array[Int] = thisInlineTensor.shape
      val padding: Float = thisInlineTensor.padding
      // (evaluating a partial sum all-reduces it: the materialised tensor is whole)
      override val distribution: Distribution =
        if (thisInlineTensor.distribution == Distribution.RowBlock) Distribution.RowBlock else Distribution.Whole
      private[compute] def doBuffer: Do[PendingBuffer] = thisInlineTensor.doBuffer
    }

    /** `Tensor::evaluate_into` of the C++ mirror: a result of at most [[DirectHostResultFloats]] floats is stored into the
      * pinned host block BY THE KERNEL (the block is device-mapped; it is wrapped as the kernel's output buffer), so the
      * read-back command disappears. The contraction pipeline is excluded: it stores through tensor maps. */
    override def flatBuffer: Do[FloatBuffer] = {
      val size = numberOfElements(shape)
      if (size == 0 || size > DirectHostResultFloats || plan.kind == 2 || distribution == Distribution.PartialSum) {
        super.flatBuffer // (the combine of a partial sum runs on a device buffer)
      } else {
        plan.arguments
          .traverse[ParallelDo, PendingBuffer](tensor => Parallel(tensor.doBuffer))
          .unwrap
          .flatMap { arguments: List[PendingBuffer] =>
            Do(TryT(ResourceT(UnitContinuation.delay {
              val hostAddress = CudaNative.hostAlloc(size * java.lang.Float.BYTES)
              Resource(value = Success(hostAddress): Try[Long], release = UnitContinuation.delay { CudaNative.hostFree(hostAddress) })
            }))).flatMap { hostAddress =>
              Do.monadicCloseable(new DeviceBuffer(CudaNative.bufferWrap(CudaNative.hostDevicePointer(hostAddress), size))).flatMap { output =>
                Do.monadicCloseable {
                    new Event(
                      CudaNative.launch(plan.handle,
                                        arguments.map(_.buffer.handle).toArray,
                                        output.handle,
                                        arguments.flatMap(_.eventOption.map(_.handle)).toArray))
                  }
                  .intransitiveFlatMap { event =>
                    Do.garbageCollected(waitForComplete(event)).map { _: Unit =>
                      MemoryUtil.memFloatBuffer(hostAddress, size.toInt)
                    }
                  }
              }
            }
          }
      }
    }
  }

  /** A constant (`Tensors.scala:1394-1397`). */
  trait FillTensor extends InlineTensor {
    def value: Float
    protected[compute] lazy val closure: FloatTerm = float.literal(value)
  }

  /** A view of `checkpoint` through an affine map (`Tensors.scala:1413-1428`). Its INPUT is a fusion barrier — `checkpoint`
    * becomes a kernel parameter — its output is not. */
  trait TransformedTensor extends InlineTensor {

    def checkpoint: Tensor

    /** number of dimensions of `checkpoint` x (number of dimensions of this view + 1), row-major */
    def matrix: MatrixData

    @transient
    protected[compute] lazy val closure: FloatTerm = checkpoint.arrayTerm.transform(matrix).extract
  }

  /** A tensor that is always a kernel parameter, never merged into a larger kernel (`Tensors.scala:1430-1440`). */
  trait NonInlineTensor extends Tensor {
    def nonInline: this.type = this

    @transient
    protected[compute] lazy val closure: FloatTerm = arrayTerm.extract
  }
}

object CudaTensors {

  private val ScalarShape: Array[Int] = Array.empty[Int]

  /** results up to this many floats are written into host memory by the kernel itself (`tensor.cpp`: kDirectHostFloats) */
  private final val DirectHostResultFloats = 16384L

  private def numberOfElements(shape: Array[Int]): Long = shape.foldLeft(1L)(_ * _)

  /** `autoBroadcastShape` (`Tensors.scala:208-222`): dimension i of the result is the non-1 (or only) one of the operands';
    * LEADING dimensions align (the opposite of NumPy), golden `TensorsSpec.scala:491-500`. */
  private def autoBroadcastShape(shape1: Array[Int], shape2: Array[Int]): Array[Int] = {
    val Absent = -1
    def dimension(shape: Array[Int], i: Int): Int = if (i < shape.length) shape(i) else Absent
    def mismatch = new IllegalArgumentException(
      s"Failed to automatically broadcast between shape [${shape1.mkString(",")}] and [${shape2.mkString(",")}]")
    Array.tabulate(math.max(shape1.length, shape2.length)) { i =>
      (dimension(shape1, i), dimension(shape2, i)) match {
        case (Absent | 1, Absent)       => throw mismatch // the reference indexes past the shorter shape here (ArrayIndexOutOfBounds)
        case (Absent | 1, other)        => other
        case (other, Absent | 1)        => other
        case (one, other) if one == other => one
        case _                          => throw mismatch
      }
    }
  }

  /** What a view (`matrix`: rows = dimensions of the row block viewed, columns = dimensions of the view + constant) of a row block is
    * (`view_of_row_block` in tensor.cpp): still a row block if the view's dimension 0 IS the block's dimension 0 and nothing else touches
    * it; a partial contribution if it fixes the block's dimension 0 to a constant (`split(0)`: the local rows, to be folded with `+`);
    * anything else would need rows of other ranks. */
  private def viewOfRowBlock(blockShape: Array[Int], viewShape: Array[Int], matrix: MatrixData): Distribution = {
    val columns = viewShape.length + 1
    val row0 = matrix.slice(0, columns)
    val row0IsIdentity = columns >= 2 && row0(0) == 1.0 && row0.tail.forall(_ == 0.0)
    val row0IsConstant = row0.init.forall(_ == 0.0)
    val othersUseDimension0 = columns >= 2 && (1 until blockShape.length).exists(row => matrix(row * columns) != 0.0)
    if (row0IsIdentity && !othersUseDimension0 && viewShape.nonEmpty && viewShape(0) == blockShape(0)) Distribution.RowBlock
    else if (row0IsConstant) Distribution.PartialSum
    else
      throw new UnsupportedOperationException(
        s"this view of a row block [${blockShape.mkString(",")}] mixes the sharded leading axis (it moves, shifts, scales or broadcasts over " +
          "dimension 0): gather it or use a replicated tensor")
  }

  /** How a tensor is spread over the ranks (one JVM per GPU) of the communicator — `Tensor::Distribution` of the C++ mirror,
    * `ct_distribution` of the C ABI. Leading-axis sharding only (SURVEY §8e). */
  sealed trait Distribution
  object Distribution {

    /** the whole tensor on every rank (also: no communicator) */
    case object Whole extends Distribution

    /** this rank's block of rows of a tensor sharded along its leading axis */
    case object RowBlock extends Distribution

    /** this rank's additive contribution to a sum over the sharded axis (the fold of `rowBlock.split(0)`) */
    case object PartialSum extends Distribution
  }

  /** The monoids `reduce` folds with (`MonoidPrograms`, `Tensors.scala:308-311`; the reference instantiates `Plus` only). */
  sealed abstract class Monoid(private[compute] val kind: Int)
  object Monoid {
    case object Plus extends Monoid(22)
    case object Min extends Monoid(20)
    case object Max extends Monoid(21)
    case object Times extends Monoid(24)
  }

  /** Kept for source compatibility with `cpu.scala:113` / `gpu.scala:25`: the hash of `Tensor.random` is part of the library's
    * precompiled `random` / `random_normal` kernels (Wang hash, `Tensors.scala:106-117`), so there is nothing to mix in. */
  trait WangHashingRandomNumberGenerator extends CudaTensors
}
