package com.thoughtworks.compute

/** Typed exceptions for the `cc_status` codes of `include/compute_cuda.h` — the counterpart of `OpenCL.Exceptions`
  * (`OpenCL.scala:143-312`), where every OpenCL error code is a case class. Illegal arguments stay the JDK's
  * `IllegalArgumentException`, which is what `Tensors.scala:208-222, 816-855` throw and `TensorsSpec.scala:66-73` expects. */
object CudaExceptions {

  sealed abstract class CudaException(val status: Int, message: String) extends RuntimeException(message)

  /** CC_ERR_NOT_INITIALIZED (-2) */
  final class NotInitialized(message: String) extends CudaException(-2, message)

  /** CC_ERR_NO_DRIVER (-3): no `libcuda.so.1` / no sm_100 device. There is no CPU fallback; cf. `DeviceNotFound` (`OpenCL.scala:160`). */
  final class DeviceNotFound(message: String) extends CudaException(-3, message)

  /** CC_ERR_CUDA (-4): a driver call failed; the message carries the `CUresult` name. */
  final class DriverError(message: String) extends CudaException(-4, message)

  /** CC_ERR_COMPILE (-5): NVRTC rejected a generated kernel; the message carries the build log like
    * `BuildProgramFailure` carries the OpenCL build logs (`OpenCL.scala:172-181, 885-915`). */
  final class BuildProgramFailure(message: String) extends CudaException(-5, message)

  /** CC_ERR_BAD_TREE (-6): malformed tree blob (a bug in [[CudaTreeWriter]], never user input). */
  final class BadTree(message: String) extends CudaException(-6, message)

  /** CC_ERR_NCCL (-7) */
  final class CollectiveError(message: String) extends CudaException(-7, message)

  /** CC_ERR_UNSUPPORTED (-8) */
  final class Unsupported(message: String) extends CudaException(-8, message)

  /** CC_ERR_OUT_OF_MEMORY (-9): `OutOfResources` / `MemObjectAllocationFailure` (`OpenCL.scala:160-170`). */
  final class OutOfResources(message: String) extends CudaException(-9, message)

  final class UnknownStatus(status: Int, message: String) extends CudaException(status, message)

  def fromStatus(status: Int, message: String): Throwable = status match {
    case -1    => new IllegalArgumentException(message)
    case -2    => new NotInitialized(message)
    case -3    => new DeviceNotFound(message)
    case -4    => new DriverError(message)
    case -5    => new BuildProgramFailure(message)
    case -6    => new BadTree(message)
    case -7    => new CollectiveError(message)
    case -8    => new Unsupported(message)
    case -9    => new OutOfResources(message)
    case other => new UnknownStatus(other, message)
  }
}
