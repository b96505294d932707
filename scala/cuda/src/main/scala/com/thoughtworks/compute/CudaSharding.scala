package com.thoughtworks.compute

/** Set-up of the communicator that sharded tensors ([[CudaTensors#Tensor.shard]], `gather`) exchange over: one JVM per GPU of
  * one box, NCCL inside `libcompute_cuda.so` plus the library's own kernels over NVLink peer memory (one-shot all-reduce /
  * all-gather of small vectors, `Tensor.sum` fused with its all-reduce, the all-gather of a row-sharded matmul fused into the
  * contraction's epilogue). The reference has nothing of the kind: it puts every device of a type into one `cl_context` and
  * leaves placement to the OpenCL runtime (`OpenCL.scala:340-374, 423-448`; SURVEY §2.4).
  *
  * Python twin (used by the tests and `bench.py` where no JVM exists): `compute/scala_b200/sharding.py`.
  *
  * {{{
  * // every rank (process), LOCAL_RANK = its GPU:
  * import com.thoughtworks.compute.cuda._
  * val uniqueId = if (rank == 0) cuda.createUniqueId() else receiveFromRank0()      // 128 bytes, shipped by any means
  * cuda.joinCommunicator(uniqueId, numberOfRanks, rank)
  * val (firstRow, rows) = cuda.rowBlock(16384, numberOfRanks, rank)
  * val x = Tensor(myRows).shard                      // this rank's [rows, 16384] block
  * x.sum                                             // the global sum: one fused kernel (local fold + all-reduce over NVLink)
  * x.split(0).reduce[Tensor](_ + _)                  // global column sums: local partial sums, all-reduced when evaluated
  * matrixMultiply(x, b).gather()                     // benchmarks.scala:188-191 on row blocks; gathered by the contraction's epilogue
  * }}}
  */
trait CudaSharding extends CudaTensors {

  /** rank 0: the 128-byte NCCL unique id to hand to the other ranks (`cc_comm_unique_id`) */
  def createUniqueId(): Array[Byte] = CudaNative.commUniqueId()

  /** Collective: `ncclCommInitRank` on this process' device, then (when `peerMemory`) one CUDA-IPC mailbox per rank mapped into every
    * other rank (`cc_comm_enable_peer`). Without NVLink peer access the small combines fall back to NCCL. */
  def joinCommunicator(uniqueId: Array[Byte], numberOfRanks: Int, rank: Int, peerMemory: Boolean = true): Unit = {
    CudaNative.commInit(uniqueId, numberOfRanks, rank)
    if (peerMemory && numberOfRanks > 1) {
      try CudaNative.commEnablePeer()
      catch {
        case _: CudaExceptions.Unsupported => // no peer access between these GPUs: NCCL carries everything
      }
    }
  }

  def leaveCommunicator(): Unit = CudaNative.commDestroy()

  /** `(numberOfRanks, rank)`; `(1, 0)` without a communicator */
  def communicator: (Int, Int) = CudaNative.commInfo()

  /** A/B switch: small combines over the NVLink peer mailboxes (default once mapped) or over NCCL */
  def routeOverPeerMemory(on: Boolean): Unit = CudaNative.commRoutePeer(on)

  /** `(first row, row count)` of `rank`'s block of a tensor with `rows` rows: the first `rows % numberOfRanks` ranks own one extra row */
  def rowBlock(rows: Int, numberOfRanks: Int, rank: Int): (Int, Int) = {
    val (first, count) = CudaNative.shardRows(rows.toLong, numberOfRanks, rank)
    (first.toInt, count.toInt)
  }

  /** B of the row-sharded matmul: `tensor` as held by `root`, on every rank (`ncclBroadcast` into a buffer of the same shape) */
  def replicate(tensor: Tensor, root: Int = 0): NonInlineTensor = {
    val replicated = tensor.nonInline
    new NonInlineTensor {
      val shape: Array[Int] = tensor.shape
      val padding: Float = tensor.padding
      private[compute] lazy val doBuffer: com.thoughtworks.raii.asynchronous.Do[PendingBuffer] = {
        import com.thoughtworks.raii.asynchronous._
        import scalaz.syntax.all._
        replicated.doBuffer.flatMap { mine =>
          // the broadcast overwrites the buffer in place: copy first, the evaluated tensor may be shared
          allocateBuffer(shape.foldLeft(1L)(_ * _)).flatMap { copy =>
            Do.monadicCloseable {
                val n = shape.foldLeft(1L)(_ * _)
                val copied = CudaNative.bufferCopy(copy.handle, mine.buffer.handle, n, mine.eventOption.map(_.handle).toArray)
                try new Event(CudaNative.broadcast(copy.handle, n, root, Array(copied)))
                finally CudaNative.eventRelease(copied)
              }
              .map(event => EventBuffer(copy, event): PendingBuffer)
          }
        }.shared
      }
    }
  }
}
