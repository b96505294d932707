package com.thoughtworks.compute

import java.nio.FloatBuffer

import com.thoughtworks.continuation._
import com.thoughtworks.future._
import com.thoughtworks.raii.asynchronous._
import com.thoughtworks.raii.covariant._
import com.thoughtworks.tryt.covariant._
import org.lwjgl.system.dyncall.DynCallback
import org.lwjgl.system.jni.JNINativeInterface
import org.lwjgl.system.{Callback, CallbackI, MemoryUtil}

import scala.concurrent.ExecutionContext
import scala.util.{Failure, Success, Try}

/** The device runtime [[CudaTensors]] sits on: the role trait `OpenCL` (`OpenCL.scala:1139-1433`) plays for `Tensors`.
  *
  * What `OpenCL` does with `cl_context` / `cl_command_queue` / `cl_mem` / `cl_event` through LWJGL, this trait does with the
  * handles of `libcompute_cuda.so` through [[CudaNative]]. The members `Tensors.scala` calls on `OpenCL` (SURVEY §8b) map as
  * follows:
  *
  *  - `allocateBuffer[Float](n)` (`OpenCL.scala:1399`)          -> [[allocateBuffer]]   (`cc_buffer_alloc`, pooled `cuMemAlloc`)
  *  - `allocateBufferFrom(hostBuffer)` (`:1415`)                -> [[allocateBufferFrom]] (`cc_buffer_from_host`, async H2D)
  *  - `DeviceBuffer.retain / release / toHostBuffer` (`:636-715`) -> [[Cuda.DeviceBuffer]]
  *  - `Event.retain / release / waitForComplete` (`:565-612`)   -> [[Cuda.Event]] (`cc_event_*`, `cc_event_on_complete`)
  *  - `createProgramWithSource` + `Program.build` + `createKernel` + `Kernel.update` + `Kernel.enqueue` + `dispatch`
  *    (`:788-844, 917-932, 1266-1329`) collapse into `cc_compile_ex` + `cc_launch`, called by `CudaTensors.enqueueClosure`;
  *    streams are picked by the library (data-hazard tracking), so `acquireCommandQueue` / `CommandQueuePool` have no counterpart;
  *  - `monadicClose` (`:1331-1337`)                             -> `cc_shutdown`.
  *
  * OpenCL-driver workarounds (`DontReleaseEventTooEarly`, `SynchronizedCreatingKernel`,
  * `HandleEventInExecutionContextForIntelAndAMDPlatform`) have no CUDA analogue and are not mirrored.
  */
trait Cuda extends MonadicCloseable[UnitContinuation] {
  import Cuda._

  /** CUDA device this backend instance drives; one process (JVM) per GPU, like `torchrun`'s `LOCAL_RANK`. */
  protected def deviceOrdinal: Int = sys.env.get("LOCAL_RANK").fold(0)(_.toInt)

  /** Replaces `numberOfCommandQueuesPerDevice` (`cpu.scala:115`, `gpu.scala:26`): compute streams commands are spread over. */
  protected def numberOfStreams: Int = 4

  protected implicit val executionContext: ExecutionContext

  CudaNative.init(deviceOrdinal)
  CudaNative.setStreamCount(numberOfStreams)

  protected type DeviceBuffer = Cuda.DeviceBuffer
  protected type Event = Cuda.Event

  /** `CommandQueue.deviceId.maxComputeUnits` (`OpenCL.scala:545`) */
  protected lazy val deviceInfo: CudaNative.DeviceInfo = CudaNative.deviceInfo()

  /** Returns an uninitialized buffer of `size` floats on the device (`OpenCL.scala:1399-1411`). */
  protected def allocateBuffer(size: Long): Do[DeviceBuffer] = Do.monadicCloseable {
    new DeviceBuffer(CudaNative.bufferAlloc(size))
  }

  /** Returns a device buffer whose content is copied from `hostBuffer` (`OpenCL.scala:1415-1431`,
    * `CL_MEM_COPY_HOST_PTR`). The copy is asynchronous: the returned event completes when `hostBuffer` may be freed. */
  protected def allocateBufferFrom(hostBuffer: FloatBuffer): Do[(DeviceBuffer, Event)] = {
    Do.delay {
      CudaNative.bufferFromHost(MemoryUtil.memAddress(hostBuffer), hostBuffer.remaining().toLong)
    }.flatMap {
      case (bufferHandle, eventHandle) =>
        Do.monadicCloseable(new DeviceBuffer(bufferHandle)).flatMap { deviceBuffer =>
          Do.monadicCloseable(new Event(eventHandle)).map { event =>
            (deviceBuffer, event)
          }
        }
    }
  }

  /** waitForStatus (`OpenCL.scala:1246-1263`): a continuation that fires on the library's callback thread once `event` completes. */
  protected def waitForComplete(event: Event): Future[Unit] = {
    val continuation: UnitContinuation[Try[Unit]] = UnitContinuation.async { (continue: Try[Unit] => Unit) =>
      val handler: Int => Unit = { status: Int =>
        // hop off the driver's callback thread before re-entering the API (the callback must not call into CUDA)
        executionContext.execute(new Runnable {
          def run(): Unit = {
            continue(if (status == 0) Success(()) else Failure(CudaExceptions.fromStatus(status, "asynchronous command failed")))
          }
        })
      }
      val userData = JNINativeInterface.NewGlobalRef(handler)
      try {
        CudaNative.eventOnComplete(event.handle, eventCallback.address(), userData)
      } catch {
        case e: Throwable =>
          JNINativeInterface.DeleteGlobalRef(userData)
          throw e
      }
    }
    Future(TryT(continuation))
  }

  /** enqueueReadBuffer + waitForComplete (`OpenCL.scala:698-715, 1206-1244`): the returned host memory is PINNED (pooled by the
    * library, `cc_host_alloc`), valid inside the `Do` scope only, exactly like the LWJGL-malloc'd buffer of the reference. */
  protected def toHostBuffer(deviceBuffer: DeviceBuffer, numberOfFloats: Int, preconditionEvents: Seq[Event]): Do[FloatBuffer] = {
    Do(TryT(ResourceT(UnitContinuation.delay {
      val hostAddress = CudaNative.hostAlloc(numberOfFloats.toLong * java.lang.Float.BYTES)
      val hostBuffer = MemoryUtil.memFloatBuffer(hostAddress, numberOfFloats)
      Resource(value = Success(hostBuffer): Try[FloatBuffer], release = UnitContinuation.delay { CudaNative.hostFree(hostAddress) })
    }))).flatMap { hostBuffer =>
      Do.monadicCloseable {
        new Event(CudaNative.bufferToHost(deviceBuffer.handle, 0L, MemoryUtil.memAddress(hostBuffer), numberOfFloats.toLong, preconditionEvents.map(_.handle).toArray))
      }.intransitiveFlatMap { event =>
        Do.garbageCollected(waitForComplete(event)).map { _: Unit =>
          hostBuffer
        }
      }
    }
  }

  /** Drops the kernel cache, the pools, the streams, the communicator and the context (`OpenCL.scala:1331-1337`). */
  def monadicClose: UnitContinuation[Unit] = UnitContinuation.delay {
    CudaNative.shutdown()
  }
}

object Cuda {

  /** A reference-counted `cc_buffer` (`DeviceBuffer`, `OpenCL.scala:636-715`): released deterministically, never by GC. */
  final class DeviceBuffer(val handle: Long) extends MonadicCloseable[UnitContinuation] {
    def retain(): Unit = CudaNative.bufferRetain(handle)
    def release(): Unit = CudaNative.bufferRelease(handle)
    def length: Long = CudaNative.bufferLength(handle)
    def monadicClose: UnitContinuation[Unit] = UnitContinuation.delay { release() }
  }

  /** A reference-counted `cc_event` (`Event`, `OpenCL.scala:565-612`). */
  final class Event(val handle: Long) extends MonadicCloseable[UnitContinuation] {
    def retain(): Unit = CudaNative.eventRetain(handle)
    def release(): Unit = CudaNative.eventRelease(handle)
    def isComplete: Boolean = CudaNative.eventQuery(handle)
    def blockingWait(): Unit = CudaNative.eventWait(handle)
    def monadicClose: UnitContinuation[Unit] = UnitContinuation.delay { release() }
  }

  /** The one native-callable function handed to `cc_event_on_complete`: `void (*)(void* user, int status)`. `user` is a JNI
    * global reference to the Scala continuation (the reference does the same for `clSetEventCallback`, `OpenCL.scala:530-537`). */
  private[compute] val eventCallback: Callback = new Callback(new CallbackI.V {
    def getSignature: String = "(pi)v"
    def callback(args: Long): Unit = {
      val userData = DynCallback.dcbArgPointer(args)
      val status = DynCallback.dcbArgInt(args)
      val handler = try MemoryUtil.memGlobalRefToObject[Int => Unit](userData)
      finally JNINativeInterface.DeleteGlobalRef(userData)
      handler(status)
    }
  }) {}

  /** Plug-in in the style of `OpenCL.GlobalExecutionContext` (`OpenCL.scala:414-416`). */
  trait GlobalExecutionContext {
    protected implicit val executionContext: ExecutionContext = ExecutionContext.global
  }
}
