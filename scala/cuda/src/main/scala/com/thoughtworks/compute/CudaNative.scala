package com.thoughtworks.compute

import java.nio.{ByteBuffer, ByteOrder}

import org.lwjgl.system.dyncall.DynCall._
import org.lwjgl.system.{Library, MemoryStack, MemoryUtil, SharedLibrary}

/** Raw binding of `include/compute_cuda.h` (libcompute_cuda.so), one method per exported `cc_*` function.
  *
  * Every entry point of the C ABI takes scalars and plain pointers and returns an `int` status, so no bespoke JNI glue is
  * needed: calls go through the generic foreign-call binding that LWJGL 3.2.3 (the version the reference pins,
  * `OpenCL/build.sbt:1-14`) ships in its core module, `org.lwjgl.system.dyncall` — the same mechanism underneath LWJGL's
  * own OpenCL binding that `OpenCL.scala` uses. From LWJGL 3.3 on replace the five `dc*` calls in [[CudaNative.Call]] with
  * `org.lwjgl.system.libffi`; nothing else changes.
  *
  * Replaces the `org.lwjgl.opencl.CL10/11/12/20` imports of `OpenCL.scala:1-20`.
  *
  * Status convention (compute_cuda.h, "Conventions"): 0 = CC_OK, negative = `cc_status`; the message of the last failure on
  * the calling thread comes from `cc_last_error()`. [[CudaNative.check]] turns it into the typed exceptions of
  * [[CudaExceptions]] the way `OpenCL.checkErrorCode` does (`OpenCL.scala:251-312`).
  */
private[compute] object CudaNative {

  /** `-Dcom.thoughtworks.compute.cuda.libname=/path/to/libcompute_cuda.so` overrides the lookup on `java.library.path`
    * (same hook LWJGL offers for the OpenCL ICD loader, `benchmarks.scala:214`). */
  private val library: SharedLibrary = {
    val explicitPath = System.getProperty("com.thoughtworks.compute.cuda.libname")
    if (explicitPath != null) Library.loadNative(explicitPath) else Library.loadNative(CudaNative.getClass, "compute_cuda")
  }

  private def address(name: String): Long = {
    val functionAddress = library.getFunctionAddress(name)
    if (functionAddress == MemoryUtil.NULL) {
      throw new UnsatisfiedLinkError(s"$name is not exported by ${library.getName}")
    }
    functionAddress
  }

  // ---- the generic caller ----------------------------------------------------------------------------------------------

  /** One dyncall VM per thread: the C ABI is thread safe and is called from arbitrary threads (`OpenCL.scala:414-416`). */
  private val callVm = new ThreadLocal[java.lang.Long] {
    override def initialValue(): java.lang.Long = {
      val vm = dcNewCallVM(4096)
      dcMode(vm, DC_CALL_C_DEFAULT)
      vm
    }
  }

  /** A C call under construction. Arguments are pushed left to right; `int32`, `int64`/`uint64`/handles, `float` and
    * pointers are the only argument types the header uses. */
  final class Call private[CudaNative] (function: Long) {
    private val vm: Long = callVm.get()
    dcReset(vm)
    def int(value: Int): Call = { dcArgInt(vm, value); this }
    def long(value: Long): Call = { dcArgLongLong(vm, value); this }
    def float(value: Float): Call = { dcArgFloat(vm, value); this }
    def pointer(value: Long): Call = { dcArgPointer(vm, value); this }
    def status(): Int = dcCallInt(vm, function)
    def checked(): Unit = check(dcCallInt(vm, function))
    def returnsPointer(): Long = dcCallPointer(vm, function)
  }
  private def call(function: Long): Call = new Call(function)

  /** checkErrorCode (`OpenCL.scala:251-312`): negative status -> typed exception carrying `cc_last_error()`. */
  def check(status: Int): Unit = {
    if (status != 0) {
      throw CudaExceptions.fromStatus(status, lastError())
    }
  }

  private def withStack[A](body: MemoryStack => A): A = {
    val stack = MemoryStack.stackPush()
    try body(stack)
    finally stack.pop()
  }

  /** Copies `values` onto the stack and returns its address, or NULL for an empty list (wait lists, argument lists). */
  private def longs(stack: MemoryStack, values: Array[Long]): Long = {
    if (values.isEmpty) MemoryUtil.NULL
    else {
      val buffer = stack.mallocLong(values.length)
      buffer.put(values)
      buffer.flip()
      MemoryUtil.memAddress(buffer)
    }
  }

  // ---- library / device (compute_cuda.h: "library / device") -------------------------------------------------------------

  private val cc_init = address("cc_init")
  private val cc_shutdown = address("cc_shutdown")
  private val cc_is_initialized = address("cc_is_initialized")
  private val cc_last_error = address("cc_last_error")
  private val cc_version = address("cc_version")
  private val cc_device_info = address("cc_device_info")
  private val cc_device_count = address("cc_device_count")
  private val cc_set_stream_count = address("cc_set_stream_count")

  /** Replaces platform / device discovery + `clCreateContext` + `CommandQueuePool` (`OpenCL.scala:340-374, 423-448, 1376-1393`). */
  def init(deviceOrdinal: Int): Unit = call(cc_init).int(deviceOrdinal).checked()

  /** monadicClose of the backend (`OpenCL.scala:1331-1337`, `Tensors.scala:1287-1289`). */
  def shutdown(): Unit = call(cc_shutdown).checked()

  def isInitialized: Boolean = call(cc_is_initialized).status() != 0

  def lastError(): String = {
    val message = call(cc_last_error).returnsPointer()
    if (message == MemoryUtil.NULL) "" else MemoryUtil.memUTF8(message)
  }

  def version(): String = MemoryUtil.memUTF8(call(cc_version).returnsPointer())

  /** `cc_device_info_t`, field offsets as declared in compute_cuda.h (natural alignment, little endian). */
  final case class DeviceInfo(ordinal: Int,
                              smCount: Int,
                              computeCapabilityMajor: Int,
                              computeCapabilityMinor: Int,
                              maxSharedMemoryPerBlock: Int,
                              l2Bytes: Int,
                              totalMemory: Long,
                              smClockKhz: Int,
                              memoryClockKhz: Int,
                              name: String)

  def deviceInfo(): DeviceInfo = withStack { stack =>
    val struct = stack.calloc(8, 112) // 6 x int32, int64, 2 x int32, char[64] = 104 bytes, rounded up to the struct's alignment
    call(cc_device_info).pointer(MemoryUtil.memAddress(struct)).checked()
    struct.order(ByteOrder.LITTLE_ENDIAN)
    val nameBytes = new Array[Byte](64)
    struct.position(40)
    struct.get(nameBytes)
    val nameLength = nameBytes.indexOf(0: Byte) match { case -1 => 64; case n => n }
    DeviceInfo(
      ordinal = struct.getInt(0),
      smCount = struct.getInt(4),
      computeCapabilityMajor = struct.getInt(8),
      computeCapabilityMinor = struct.getInt(12),
      maxSharedMemoryPerBlock = struct.getInt(16),
      l2Bytes = struct.getInt(20),
      totalMemory = struct.getLong(24),
      smClockKhz = struct.getInt(32),
      memoryClockKhz = struct.getInt(36),
      name = new String(nameBytes, 0, nameLength, "UTF-8")
    )
  }

  def deviceCount(): Int = withStack { stack =>
    val out = stack.mallocInt(1)
    call(cc_device_count).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0)
  }

  /** Replaces `numberOfCommandQueuesPerDevice` (`cpu.scala:115`). */
  def setStreamCount(numberOfStreams: Int): Unit = call(cc_set_stream_count).int(numberOfStreams).checked()

  // ---- memory (compute_cuda.h: "memory") ----------------------------------------------------------------------------------

  private val cc_buffer_alloc = address("cc_buffer_alloc")
  private val cc_buffer_from_host = address("cc_buffer_from_host")
  private val cc_buffer_upload = address("cc_buffer_upload")
  private val cc_buffer_wrap = address("cc_buffer_wrap")
  private val cc_buffer_retain = address("cc_buffer_retain")
  private val cc_buffer_release = address("cc_buffer_release")
  private val cc_buffer_device_ptr = address("cc_buffer_device_ptr")
  private val cc_buffer_length = address("cc_buffer_length")
  private val cc_buffer_to_host = address("cc_buffer_to_host")
  private val cc_memory_trim = address("cc_memory_trim")
  private val cc_host_alloc = address("cc_host_alloc")
  private val cc_host_free = address("cc_host_free")
  private val cc_host_device_ptr = address("cc_host_device_ptr")

  private def outHandle(stack: MemoryStack)(invoke: Long => Unit): Long = {
    val out = stack.mallocLong(1)
    out.put(0, 0L)
    invoke(MemoryUtil.memAddress(out))
    out.get(0)
  }

  /** allocateBuffer[Float](n) (`OpenCL.scala:1399-1411`). */
  def bufferAlloc(numberOfFloats: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_buffer_alloc).long(numberOfFloats).pointer(out).checked())
  }

  /** allocateBufferFrom(hostBuffer) (`OpenCL.scala:1415-1431`): asynchronous H2D. Returns `(buffer, event)`; the host memory
    * must stay valid until the event completes. */
  def bufferFromHost(hostAddress: Long, numberOfFloats: Long): (Long, Long) = withStack { stack =>
    val out = stack.mallocLong(2)
    out.put(0, 0L).put(1, 0L)
    val base = MemoryUtil.memAddress(out)
    call(cc_buffer_from_host).pointer(hostAddress).long(numberOfFloats).pointer(base).pointer(base + 8).checked()
    (out.get(0), out.get(1))
  }

  def bufferUpload(buffer: Long, hostAddress: Long, numberOfFloats: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_buffer_upload).long(buffer).pointer(hostAddress).long(numberOfFloats).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  def bufferWrap(devicePointer: Long, numberOfFloats: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_buffer_wrap).long(devicePointer).long(numberOfFloats).pointer(out).checked())
  }

  /** DeviceBuffer.retain / release (`OpenCL.scala:644-648`). */
  def bufferRetain(buffer: Long): Unit = call(cc_buffer_retain).long(buffer).checked()
  def bufferRelease(buffer: Long): Unit = call(cc_buffer_release).long(buffer).checked()

  def bufferDevicePointer(buffer: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_buffer_device_ptr).long(buffer).pointer(out).checked())
  }

  def bufferLength(buffer: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_buffer_length).long(buffer).pointer(out).checked())
  }

  /** enqueueReadBuffer (`OpenCL.scala:1206-1244`): asynchronous D2H after `waits`; returns the completion event. */
  def bufferToHost(buffer: Long, offsetInFloats: Long, hostAddress: Long, numberOfFloats: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_buffer_to_host).long(buffer).long(offsetInFloats).pointer(hostAddress).long(numberOfFloats).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  def memoryTrim(): Unit = call(cc_memory_trim).checked()

  /** Pinned, device-mapped, pooled host memory; replaces LWJGL `memAllocFloat` (`Memory.scala:184-208`). */
  def hostAlloc(numberOfBytes: Long): Long = withStack { stack =>
    val out = stack.mallocPointer(1)
    call(cc_host_alloc).long(numberOfBytes).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0)
  }
  def hostFree(hostAddress: Long): Unit = call(cc_host_free).pointer(hostAddress).checked()
  def hostDevicePointer(hostAddress: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_host_device_ptr).pointer(hostAddress).pointer(out).checked())
  }

  // ---- events (compute_cuda.h: "events") -------------------------------------------------------------------------------------

  private val cc_event_retain = address("cc_event_retain")
  private val cc_event_release = address("cc_event_release")
  private val cc_event_wait = address("cc_event_wait")
  private val cc_event_query = address("cc_event_query")
  private val cc_event_on_complete = address("cc_event_on_complete")
  private val cc_synchronize = address("cc_synchronize")

  def eventRetain(event: Long): Unit = call(cc_event_retain).long(event).checked()
  def eventRelease(event: Long): Unit = call(cc_event_release).long(event).checked()
  def eventWait(event: Long): Unit = call(cc_event_wait).long(event).checked()
  def eventQuery(event: Long): Boolean = withStack { stack =>
    val out = stack.mallocInt(1)
    call(cc_event_query).long(event).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0) != 0
  }

  /** clSetEventCallback replacement (`OpenCL.scala:1246-1263`): `callback` is the address of a native-callable function
    * `void (*)(void* user, int status)` (see [[Cuda.eventCallback]]), `userData` is handed back to it. */
  def eventOnComplete(event: Long, callback: Long, userData: Long): Unit =
    call(cc_event_on_complete).long(event).pointer(callback).pointer(userData).checked()

  def synchronize(): Unit = call(cc_synchronize).checked()

  // ---- expression trees -> kernels (compute_cuda.h: "expression trees -> kernels") ---------------------------------------------

  private val cc_compile = address("cc_compile")
  private val cc_compile_ex = address("cc_compile_ex")
  private val cc_kernel_disk_cache = address("cc_kernel_disk_cache")
  private val cc_kernel_cache_limit = address("cc_kernel_cache_limit")
  private val cc_kernel_cache_clear = address("cc_kernel_cache_clear")
  private val cc_kernel_cache_size = address("cc_kernel_cache_size")
  private val cc_kernel_cache_lookup = address("cc_kernel_cache_lookup")
  private val cc_kernel_retain = address("cc_kernel_retain")
  private val cc_kernel_release = address("cc_kernel_release")
  private val cc_kernel_info = address("cc_kernel_info")
  private val cc_kernel_arg_param = address("cc_kernel_arg_param")
  private val cc_kernel_source = address("cc_kernel_source")
  private val cc_launch = address("cc_launch")

  /** createProgramWithSource + build + cache probe (`Tensors.scala:1293-1331`) in one call. `treeBlob` must be a direct
    * buffer positioned at the blob ([[CudaTreeWriter.finish]]). */
  def compile(treeBlob: ByteBuffer): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_compile).pointer(MemoryUtil.memAddress(treeBlob)).long(treeBlob.remaining().toLong).pointer(out).checked())
  }

  /** Same, and reports the blob's parameter ids in ordinal order (identity-deduplicated DFS pre-order of the main tree =
    * `parameterDescendants`, `Tensors.scala:230-251`, then the parameters first met inside definitions). */
  def compileEx(treeBlob: ByteBuffer, capacity: Int): (Long, Array[Long]) = withStack { stack =>
    val ids = stack.mallocLong(math.max(capacity, 1))
    val count = stack.mallocInt(1)
    val kernel = outHandle(stack) { out =>
      call(cc_compile_ex)
        .pointer(MemoryUtil.memAddress(treeBlob))
        .long(treeBlob.remaining().toLong)
        .pointer(out)
        .pointer(MemoryUtil.memAddress(ids))
        .int(capacity)
        .pointer(MemoryUtil.memAddress(count))
        .checked()
    }
    (kernel, Array.tabulate(count.get(0))(ids.get(_)))
  }

  def kernelDiskCache(directory: String): Unit = withStack { stack =>
    call(cc_kernel_disk_cache).pointer(if (directory == null) MemoryUtil.NULL else MemoryUtil.memAddress(stack.UTF8(directory))).checked()
  }
  def kernelCacheLimit(maximumNumberOfKernels: Long): Unit = call(cc_kernel_cache_limit).long(maximumNumberOfKernels).checked()
  def kernelCacheClear(): Unit = call(cc_kernel_cache_clear).checked()
  def kernelCacheSize(): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_kernel_cache_size).pointer(out).checked())
  }

  /** Probe only (never compiles): the cached kernel with the blob's structure — `kernelCache.getIfPresent`
    * (`Tensors.scala:1293`, `TensorsSpec.scala:50-52`) — retained for the caller, or 0. */
  def kernelCacheLookup(treeBlob: ByteBuffer, anyOutputShape: Boolean): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_kernel_cache_lookup).pointer(MemoryUtil.memAddress(treeBlob)).long(treeBlob.remaining().toLong).int(if (anyOutputShape) 1 else 0).pointer(out).checked()
    }
  }
  def kernelRetain(kernel: Long): Unit = call(cc_kernel_retain).long(kernel).checked()
  def kernelRelease(kernel: Long): Unit = call(cc_kernel_release).long(kernel).checked()

  /** `cc_kernel_info_t`. `kind`: 0 elementwise, 1 axis reduction, 2 contraction (tcgen05), 3 tiled transpose, 4 whole-tensor fold. */
  final case class KernelInfo(kind: Int,
                              cacheHit: Boolean,
                              numberOfArguments: Int,
                              numberOfLaunches: Int,
                              outputFloats: Long,
                              algorithmicBytes: Long,
                              flops: Long,
                              structuralHash: Long)

  def kernelInfo(kernel: Long): KernelInfo = withStack { stack =>
    val struct = stack.calloc(8, 48) // 4 x int32, 4 x uint64
    call(cc_kernel_info).long(kernel).pointer(MemoryUtil.memAddress(struct)).checked()
    struct.order(ByteOrder.LITTLE_ENDIAN)
    KernelInfo(struct.getInt(0), struct.getInt(4) != 0, struct.getInt(8), struct.getInt(12), struct.getLong(16), struct.getLong(24), struct.getLong(32), struct.getLong(40))
  }

  /** which tree parameter (ordinal, see [[compileEx]]) the i-th buffer argument of `cc_launch` is */
  def kernelArgumentParameter(kernel: Long, argumentIndex: Int): Int = withStack { stack =>
    val out = stack.mallocInt(1)
    call(cc_kernel_arg_param).long(kernel).int(argumentIndex).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0)
  }

  def kernelSource(kernel: Long): String = withStack { stack =>
    val out = stack.mallocPointer(1)
    call(cc_kernel_source).long(kernel).pointer(MemoryUtil.memAddress(out)).checked()
    MemoryUtil.memUTF8(out.get(0))
  }

  /** Kernel.enqueue + dispatch (`OpenCL.scala:788-844, 1298-1329`; `Tensors.scala:1342-1375`): returns the completion event
    * (already retained for the caller). */
  def launch(kernel: Long, arguments: Array[Long], output: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_launch).long(kernel).pointer(longs(stack, arguments)).int(arguments.length).long(output).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  // ---- precompiled programs: Tensor.sum, random, randomNormal, the contraction (`Tensors.scala:303-443, 673-771`) -----------------------

  private val cc_reduce_sum = address("cc_reduce_sum")
  private val cc_random = address("cc_random")
  private val cc_random_normal = address("cc_random_normal")
  private val cc_matmul_3xtf32 = address("cc_matmul_3xtf32")
  private val cc_set_operand_cache = address("cc_set_operand_cache")

  def reduceSum(input: Long, numberOfFloats: Long, output: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_reduce_sum).long(input).long(numberOfFloats).long(output).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  def random(output: Long, numberOfFloats: Long, seed: Int): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_random).long(output).long(numberOfFloats).int(seed).pointer(out).checked())
  }

  def randomNormal(output: Long, numberOfFloats: Long, seed: Int): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_random_normal).long(output).long(numberOfFloats).int(seed).pointer(out).checked())
  }

  def matmul3xTf32(a: Long, b: Long, c: Long, m: Long, n: Long, k: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_matmul_3xtf32).long(a).long(b).long(c).long(m).long(n).long(k).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  def setOperandCache(on: Boolean): Unit = call(cc_set_operand_cache).int(if (on) 1 else 0).checked()

  // ---- counters / timing ------------------------------------------------------------------------------------------------------------

  private val cc_stats = address("cc_stats")
  private val cc_stats_reset = address("cc_stats_reset")
  private val cc_profile_enable = address("cc_profile_enable")
  private val cc_profile_report = address("cc_profile_report")
  private val cc_timer_start = address("cc_timer_start")
  private val cc_timer_stop = address("cc_timer_stop")

  /** `cc_stats_t`: twelve uint64 counters in declaration order. */
  final case class Stats(compiles: Long,
                         cacheHits: Long,
                         launches: Long,
                         deviceKernels: Long,
                         hostToDeviceBytes: Long,
                         deviceToHostBytes: Long,
                         allocCalls: Long,
                         poolHits: Long,
                         bytesInUse: Long,
                         bytesPooled: Long,
                         nvrtcCompiles: Long,
                         diskCacheHits: Long)

  def stats(): Stats = withStack { stack =>
    val struct = stack.callocLong(12)
    call(cc_stats).pointer(MemoryUtil.memAddress(struct)).checked()
    Stats(struct.get(0), struct.get(1), struct.get(2), struct.get(3), struct.get(4), struct.get(5), struct.get(6), struct.get(7), struct.get(8), struct.get(9), struct.get(10), struct.get(11))
  }
  def statsReset(): Unit = call(cc_stats_reset).checked()
  def profileEnable(on: Boolean): Unit = call(cc_profile_enable).int(if (on) 1 else 0).checked()

  /** JSON array, one record per kernel structure / copy direction / collective. */
  def profileReport(): String = withStack { stack =>
    val needed = stack.mallocLong(1)
    call(cc_profile_report).pointer(MemoryUtil.NULL).long(0L).pointer(MemoryUtil.memAddress(needed)).checked()
    val buffer = MemoryUtil.memAlloc(needed.get(0).toInt + 1)
    try {
      call(cc_profile_report).pointer(MemoryUtil.memAddress(buffer)).long(buffer.capacity().toLong).pointer(MemoryUtil.memAddress(needed)).checked()
      MemoryUtil.memUTF8(MemoryUtil.memAddress(buffer))
    } finally MemoryUtil.memFree(buffer)
  }
  def timerStart(): Unit = call(cc_timer_start).checked()
  def timerStopMilliseconds(): Float = withStack { stack =>
    val out = stack.mallocFloat(1)
    call(cc_timer_stop).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0)
  }

  // ---- multi-GPU: one JVM per GPU, NCCL + our own kernels over NVLink peer memory ------------------------------------------------------

  private val cc_comm_unique_id = address("cc_comm_unique_id")
  private val cc_comm_init = address("cc_comm_init")
  private val cc_comm_destroy = address("cc_comm_destroy")
  private val cc_comm_info = address("cc_comm_info")
  private val cc_comm_enable_peer = address("cc_comm_enable_peer")
  private val cc_comm_peer_enabled = address("cc_comm_peer_enabled")
  private val cc_comm_route_peer = address("cc_comm_route_peer")
  private val cc_reduce_sum_allreduce = address("cc_reduce_sum_allreduce")
  private val cc_allreduce_sum = address("cc_allreduce_sum")
  private val cc_comm_symmetric_alloc = address("cc_comm_symmetric_alloc")
  private val cc_matmul_3xtf32_allgather = address("cc_matmul_3xtf32_allgather")
  private val cc_allgather = address("cc_allgather")
  private val cc_broadcast = address("cc_broadcast")

  /** rank 0 only; ship the 128 bytes to the other JVMs out of band */
  def commUniqueId(): Array[Byte] = withStack { stack =>
    val out = stack.malloc(128)
    call(cc_comm_unique_id).pointer(MemoryUtil.memAddress(out)).checked()
    val bytes = new Array[Byte](128)
    out.get(bytes)
    bytes
  }
  def commInit(uniqueId: Array[Byte], numberOfRanks: Int, rank: Int): Unit = withStack { stack =>
    require(uniqueId.length == 128)
    val in = stack.malloc(128)
    in.put(uniqueId).flip()
    call(cc_comm_init).pointer(MemoryUtil.memAddress(in)).int(numberOfRanks).int(rank).checked()
  }
  def commDestroy(): Unit = call(cc_comm_destroy).checked()
  def commInfo(): (Int, Int) = withStack { stack =>
    val out = stack.mallocInt(2)
    val base = MemoryUtil.memAddress(out)
    call(cc_comm_info).pointer(base).pointer(base + 4).checked()
    (out.get(0), out.get(1))
  }
  def commEnablePeer(): Unit = call(cc_comm_enable_peer).checked()
  def commPeerEnabled(): Boolean = withStack { stack =>
    val out = stack.mallocInt(1)
    call(cc_comm_peer_enabled).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0) != 0
  }
  def commRoutePeer(on: Boolean): Unit = call(cc_comm_route_peer).int(if (on) 1 else 0).checked()
  def reduceSumAllReduce(input: Long, numberOfFloats: Long, output: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_reduce_sum_allreduce).long(input).long(numberOfFloats).long(output).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }
  def allReduceSum(buffer: Long, numberOfFloats: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_allreduce_sum).long(buffer).long(numberOfFloats).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked())
  }
  def commSymmetricAlloc(numberOfFloats: Long): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_comm_symmetric_alloc).long(numberOfFloats).pointer(out).checked())
  }
  def matmul3xTf32AllGather(aShard: Long, b: Long, gathered: Long, mShard: Long, n: Long, k: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_matmul_3xtf32_allgather).long(aShard).long(b).long(gathered).long(mShard).long(n).long(k).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }
  def allGather(send: Long, receive: Long, numberOfFloatsPerRank: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_allgather).long(send).long(receive).long(numberOfFloatsPerRank).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }
  def broadcast(buffer: Long, numberOfFloats: Long, root: Int, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_broadcast).long(buffer).long(numberOfFloats).int(root).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  def commGeneration(): Long = withStack { stack =>
    outHandle(stack)(out => call(cc_comm_generation).pointer(out).checked())
  }

  // ---- leading-axis sharding (compute_cuda.h: "leading-axis sharding over the GPUs of one box") ----------------------------------------

  private val cc_comm_generation = address("cc_comm_generation")
  private val cc_buffer_copy = address("cc_buffer_copy")
  private val cc_shard_rows = address("cc_shard_rows")
  private val cc_shard_agree = address("cc_shard_agree")
  private val cc_shard_launch_allreduce = address("cc_shard_launch_allreduce")
  private val cc_shard_launch_allgather = address("cc_shard_launch_allgather")

  def bufferCopy(destination: Long, source: Long, numberOfFloats: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_buffer_copy).long(destination).long(source).long(numberOfFloats).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  /** `(first row, row count)` of `rank`'s block: the first `rows % numberOfRanks` ranks own one extra row */
  def shardRows(rows: Long, numberOfRanks: Int, rank: Int): (Long, Long) = withStack { stack =>
    val out = stack.mallocLong(2)
    val base = MemoryUtil.memAddress(out)
    call(cc_shard_rows).long(rows).int(numberOfRanks).int(rank).pointer(base).pointer(base + 8).checked()
    (out.get(0), out.get(1))
  }

  /** collective: did every rank pass the same value? */
  def shardAgree(value: Long): Boolean = withStack { stack =>
    val out = stack.mallocInt(1)
    call(cc_shard_agree).long(value).pointer(MemoryUtil.memAddress(out)).checked()
    out.get(0) != 0
  }

  /** `cc_launch`, then all-reduce (sum) of the output across ranks, in place */
  def shardLaunchAllReduce(kernel: Long, arguments: Array[Long], output: Long, waits: Array[Long]): Long = withStack { stack =>
    outHandle(stack) { out =>
      call(cc_shard_launch_allreduce).long(kernel).pointer(longs(stack, arguments)).int(arguments.length).long(output).pointer(longs(stack, waits)).int(waits.length).pointer(out).checked()
    }
  }

  /** `cc_launch` on this rank's row block with the result gathered on every rank; returns `(event, fused)` — `fused` tells whether
    * the exchange ran inside the contraction's epilogue */
  def shardLaunchAllGather(kernel: Long, arguments: Array[Long], gathered: Long, waits: Array[Long]): (Long, Boolean) = withStack { stack =>
    val fused = stack.mallocInt(1)
    val event = outHandle(stack) { out =>
      call(cc_shard_launch_allgather)
        .long(kernel)
        .pointer(longs(stack, arguments))
        .int(arguments.length)
        .long(gathered)
        .pointer(longs(stack, waits))
        .int(waits.length)
        .pointer(out)
        .pointer(MemoryUtil.memAddress(fused))
        .checked()
    }
    (event, fused.get(0) != 0)
  }
}
