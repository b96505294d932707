package com.thoughtworks.compute

import com.typesafe.scalalogging.StrictLogging

/** Contains N-dimensional array types on NVIDIA B200 GPUs (sm_100a), backed by `libcompute_cuda.so`.
  *
  * All the usage of this [[cuda]] object is the same as [[cpu]] (`cpu.scala:15-117`) and [[gpu]] (`gpu.scala:15-27`), except
  * the `import` statement:
  *
  * {{{
  * import com.thoughtworks.compute.cuda._
  *
  * val a = Tensor(Array(Seq(1.0f, 2.0f, 3.0f), Seq(4.0f, 5.0f, 6.0f)))
  * Tensor.tanh(a * a + a).toString
  * }}}
  *
  * What differs underneath: expression graphs are serialised ([[CudaTreeWriter]]) and compiled once per structure by a CUDA
  * code generator + NVRTC for sm_100a behind a C ABI ([[CudaNative]]); per-axis sums, the matmul formulations of
  * `benchmarks.scala:174-193` and the convolution of `benchmarks.scala:463-556` — which users write as unrolled chains over
  * `split` — are re-rolled into real reductions and a TMA + tcgen05 (3xTF32) contraction. There is no CPU fallback: without a
  * CUDA driver and an sm_100 device the object fails to initialise with [[CudaExceptions.DeviceNotFound]].
  *
  * One JVM per GPU: the device is `LOCAL_RANK` (default 0). See [[CudaSharding]] for tensors sharded over the GPUs of one box.
  */
object cuda
    extends StrictLogging
    with Cuda.GlobalExecutionContext
    with CudaTensors.WangHashingRandomNumberGenerator
    with CudaSharding {

  /** the counterpart of `numberOfCommandQueuesPerDevice = 5` (`cpu.scala:115`, `gpu.scala:26`) */
  override protected val numberOfStreams: Int = 4
}
