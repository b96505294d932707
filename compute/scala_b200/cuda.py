"""`com.thoughtworks.compute.cuda` — the new backend object beside `cpu` / `gpu` (cpu.scala:103-117, gpu.scala:15-27),
as seen from Python.  This module is a ctypes VIEW of the host-side mirror compiled into libcompute_cuda.so
(csrc/tensor.cpp): names, argument meaning and error behaviour follow Tensors.scala:395-1442; nothing is computed
here and nothing falls back to numpy.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from ._lib import ComputeCudaError, IllegalArgumentException, check, i32, u64  # noqa: F401

_UNARY = {"exp": 10, "log": 11, "abs": 12, "tanh": 13, "sqrt": 14, "neg": 15}
_BINARY = {"min": 20, "max": 21, "+": 22, "-": 23, "*": 24, "/": 25, "%": 26}

_configured = False


def _L():
    global _configured
    L = _lib.lib()
    if not _configured:
        f32, h, ip, fp = C.c_float, u64, C.POINTER(i32), C.POINTER(C.c_float)
        hp = C.POINTER(u64)
        L.ct_from_host.argtypes = [C.c_void_p, ip, C.c_int, f32, hp]
        L.ct_from_buffer.argtypes = [h, ip, C.c_int, f32, hp]
        L.ct_scalar.argtypes = [f32, f32, hp]
        L.ct_fill.argtypes = [f32, ip, C.c_int, f32, hp]
        L.ct_random.argtypes = [ip, C.c_int, i32, f32, hp]
        L.ct_random_normal.argtypes = [ip, C.c_int, i32, f32, hp]
        L.ct_unary.argtypes = [C.c_int, h, hp]
        L.ct_binary.argtypes = [C.c_int, h, h, hp]
        for n in ("ct_broadcast", "ct_reshape", "ct_scale"):
            getattr(L, n).argtypes = [h, ip, C.c_int, hp]
        L.ct_translate.argtypes = [h, C.POINTER(C.c_double), C.c_int, ip, C.c_int, hp]
        L.ct_permute.argtypes = [h, ip, C.c_int, hp]
        L.ct_transpose.argtypes = [h, hp]
        L.ct_split.argtypes = [h, C.c_int, hp, C.c_int, C.POINTER(C.c_int)]
        L.ct_join.argtypes = [hp, C.c_int, hp]
        L.ct_join_dim.argtypes = [hp, C.c_int, C.c_int, hp]
        for n in ("ct_sum", "ct_non_inline", "ct_do_cache"):
            getattr(L, n).argtypes = [h, hp]
        L.ct_reduce.argtypes = [h, C.c_int, hp]
        L.ct_rank.argtypes = [h, C.POINTER(C.c_int)]
        L.ct_shape.argtypes = [h, ip, C.c_int]
        L.ct_padding.argtypes = [h, fp]
        L.ct_flat_array.argtypes = [h, C.c_void_p, u64]
        L.ct_flat_buffer.argtypes = [h, C.POINTER(C.POINTER(C.c_float)), hp]
        L.ct_flat_buffer_release.argtypes = [C.POINTER(C.c_float)]
        L.ct_to_string.argtypes = [h, C.c_char_p, u64, hp]
        L.ct_do_buffer.argtypes = [h, hp, hp]
        L.ct_compile.argtypes = [h, hp]
        L.ct_release.argtypes = [h]
        L.ct_live_tensors.argtypes = [C.POINTER(C.c_int64)]
        L.cc_profile_enable.argtypes = [C.c_int]
        L.cc_profile_report.argtypes = [C.c_char_p, u64, hp]
        L.cc_kernel_disk_cache.argtypes = [C.c_char_p]
        L.cc_init.argtypes = [C.c_int]
        L.cc_set_stream_count.argtypes = [C.c_int]
        L.cc_device_info.argtypes = [C.POINTER(_lib.DeviceInfo)]
        L.cc_buffer_alloc.argtypes = [u64, hp]
        L.cc_buffer_from_host.argtypes = [C.c_void_p, u64, hp, hp]
        L.cc_buffer_upload.argtypes = [h, C.c_void_p, u64, hp, C.c_int, hp]
        L.cc_buffer_wrap.argtypes = [u64, u64, hp]
        L.cc_buffer_retain.argtypes = [h]
        L.cc_buffer_release.argtypes = [h]
        L.cc_buffer_device_ptr.argtypes = [h, hp]
        L.cc_buffer_length.argtypes = [h, hp]
        L.cc_buffer_to_host.argtypes = [h, u64, C.c_void_p, u64, hp, C.c_int, hp]
        L.cc_host_alloc.argtypes = [u64, C.POINTER(C.c_void_p)]
        L.cc_host_free.argtypes = [C.c_void_p]
        L.cc_host_device_ptr.argtypes = [C.c_void_p, hp]
        for n in ("cc_event_retain", "cc_event_release", "cc_event_wait"):
            getattr(L, n).argtypes = [h]
        L.cc_event_query.argtypes = [h, C.POINTER(C.c_int)]
        L.cc_event_on_complete.argtypes = [h, C.c_void_p, C.c_void_p]
        L.cc_compile.argtypes = [C.c_void_p, u64, hp]
        L.cc_compile_ex.argtypes = [C.c_void_p, u64, hp, hp, C.c_int, C.POINTER(C.c_int)]
        L.cc_kernel_retain.argtypes = [h]
        L.cc_kernel_release.argtypes = [h]
        L.cc_kernel_info.argtypes = [h, C.POINTER(_lib.KernelInfo)]
        L.cc_kernel_arg_param.argtypes = [h, C.c_int, ip]
        L.cc_kernel_source.argtypes = [h, C.POINTER(C.c_char_p)]
        L.cc_launch.argtypes = [h, hp, C.c_int, h, hp, C.c_int, hp]
        L.cc_reduce_sum.argtypes = [h, u64, h, hp, C.c_int, hp]
        L.cc_random.argtypes = [h, u64, i32, hp]
        L.cc_random_normal.argtypes = [h, u64, i32, hp]
        L.cc_matmul_3xtf32.argtypes = [h, h, h, C.c_int64, C.c_int64, C.c_int64, hp, C.c_int, hp]
        L.cc_set_operand_cache.argtypes = [C.c_int]
        L.cc_kernel_cache_limit.argtypes = [u64]
        L.cc_kernel_cache_size.argtypes = [hp]
        L.cc_comm_symmetric_alloc.argtypes = [u64, hp]
        L.cc_matmul_3xtf32_allgather.argtypes = [h, h, h, C.c_int64, C.c_int64, C.c_int64, hp, C.c_int, hp]
        L.cc_stats.argtypes = [C.POINTER(_lib.Stats)]
        L.cc_timer_stop.argtypes = [fp]
        L.cc_comm_unique_id.argtypes = [C.c_void_p]
        L.cc_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cc_comm_info.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cc_allreduce_sum.argtypes = [h, u64, hp, C.c_int, hp]
        L.cc_reduce_sum_allreduce.argtypes = [h, u64, h, hp, C.c_int, hp]
        L.cc_comm_peer_enabled.argtypes = [C.POINTER(C.c_int)]
        L.cc_comm_route_peer.argtypes = [C.c_int]
        L.cc_allgather.argtypes = [h, h, u64, hp, C.c_int, hp]
        L.cc_broadcast.argtypes = [h, u64, C.c_int, hp, C.c_int, hp]
        L.cc_buffer_copy.argtypes = [h, h, u64, hp, C.c_int, hp]
        L.cc_buffer_is_multicast.argtypes = [h, C.POINTER(C.c_int)]
        L.cc_kernel_cache_lookup.argtypes = [C.c_void_p, u64, C.c_int, hp]
        L.cc_comm_generation.argtypes = [hp]
        L.cc_shard_rows.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.cc_shard_agree.argtypes = [u64, C.POINTER(C.c_int)]
        L.cc_shard_launch_allreduce.argtypes = [h, hp, C.c_int, h, hp, C.c_int, hp]
        L.cc_shard_launch_allgather.argtypes = [h, hp, C.c_int, h, hp, C.c_int, hp, C.POINTER(C.c_int)]
        L.cc_graph_end.argtypes = [hp]
        L.cc_graph_launch.argtypes = [h, hp, C.c_int, hp]
        L.cc_graph_info.argtypes = [h, hp, hp]
        L.cc_graph_release.argtypes = [h]
        L.ct_tree_blob.argtypes = [h, C.c_void_p, u64, hp]
        L.ct_shard.argtypes = [h, hp]
        L.ct_distribution.argtypes = [h, C.POINTER(C.c_int)]
        L.ct_gather.argtypes = [h, C.c_int, hp]
        _configured = True
    return L


def _shape_arg(shape: Sequence[int]):
    shape = [int(s) for s in shape]
    return (i32 * max(1, len(shape)))(*shape), len(shape)


# ---- runtime (trait OpenCL's role) ----------------------------------------------------------------------------------------


def init(device: int = -1, streams: int | None = None) -> None:
    """Factory[... cuda ...].newInstance(): creates the context / stream pool. Raises without a B200 + driver."""
    L = _L()
    if streams is not None and not L.cc_is_initialized():
        check(L.cc_set_stream_count(int(streams)))
    check(L.cc_init(int(device)))


def shutdown() -> None:
    check(_L().cc_shutdown())


def is_initialized() -> bool:
    return bool(_L().cc_is_initialized())


def device_info() -> _lib.DeviceInfo:
    info = _lib.DeviceInfo()
    check(_L().cc_device_info(C.byref(info)))
    return info


def stats() -> dict:
    s = _lib.Stats()
    check(_L().cc_stats(C.byref(s)))
    return {n: int(getattr(s, n)) for n, _ in s._fields_}


def memory_trim() -> None:
    """return idle pooled device memory to the driver"""
    check(_L().cc_memory_trim())


def stats_reset() -> None:
    check(_L().cc_stats_reset())


def profile(on: bool = True) -> None:
    """built-in command profiler: while on, every kernel launch / copy / collective is timed on the device"""
    check(_L().cc_profile_enable(1 if on else 0))


def profile_report() -> list:
    """device time per kernel structure / copy direction / collective since the last report (consumes the records)"""
    import json

    need = u64()
    check(_L().cc_profile_report(None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    check(_L().cc_profile_report(buf, need.value, None))
    return json.loads(buf.value.decode())


def synchronize() -> None:
    check(_L().cc_synchronize())


def timer_start() -> None:
    check(_L().cc_timer_start())


def timer_stop() -> float:
    ms = C.c_float()
    check(_L().cc_timer_stop(C.byref(ms)))
    return float(ms.value)


class PinnedArray:
    """pinned host staging memory (cc_host_alloc) exposed as a numpy float32 array"""

    def __init__(self, n_floats: int):
        self._p = C.c_void_p()
        self.n = int(n_floats)
        check(_L().cc_host_alloc(self.n * 4, C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_float)), shape=(max(self.n, 1),))[: self.n]

    @property
    def ptr(self) -> int:
        return self._p.value

    def free(self) -> None:
        if self._p:
            self.array = None
            check(_L().cc_host_free(self._p))
            self._p = C.c_void_p()


class HostBuffer:
    """`flatBuffer`'s result (T:1099-1109): callee-allocated pinned host memory holding the evaluated tensor, valid until
    `release()` (the end of the reference's `Do` scope, O:691-715).  Usable as a context manager."""

    __slots__ = ("_p", "n", "array")

    def __init__(self, ptr: int, view):
        self._p = ptr
        self.array = np.frombuffer(view, dtype=np.float32) if len(view) else np.empty(0, dtype=np.float32)
        self.n = self.array.size

    def release(self) -> None:
        if self._p is not None:
            self.array = None
            p, self._p = self._p, None
            st = (_HOT.flat_buffer_release or _hot().flat_buffer_release)(p)
            if st:
                check(st)

    def __enter__(self) -> np.ndarray:
        return self.array

    def __exit__(self, *exc) -> None:
        self.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Kernel:
    """CompiledKernel (Tensors.scala:1263-1265) handle, for cache / pattern tests"""

    def __init__(self, handle: int):
        self.handle = handle

    @property
    def info(self) -> _lib.KernelInfo:
        k = _lib.KernelInfo()
        check(_L().cc_kernel_info(self.handle, C.byref(k)))
        return k

    @property
    def source(self) -> str:
        s = C.c_char_p()
        check(_L().cc_kernel_source(self.handle, C.byref(s)))
        return s.value.decode()

    def release(self) -> None:
        if self.handle:
            check(_L().cc_kernel_release(self.handle))
            self.handle = 0

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Buffer:
    """DeviceBuffer[Float] handle (OpenCL.scala:636-715)"""

    __slots__ = ("handle",)

    def __init__(self, handle: int):
        self.handle = handle

    @staticmethod
    def alloc(n_floats: int) -> "Buffer":
        h = u64()
        check(_L().cc_buffer_alloc(int(n_floats), C.byref(h)))
        return Buffer(h.value)

    @staticmethod
    def wrap(device_ptr: int, n_floats: int) -> "Buffer":
        h = u64()
        check(_L().cc_buffer_wrap(int(device_ptr), int(n_floats), C.byref(h)))
        return Buffer(h.value)

    @staticmethod
    def from_host(a: np.ndarray) -> "Buffer":
        a = np.ascontiguousarray(a, dtype=np.float32)
        h = u64()
        check(_L().cc_buffer_from_host(a.ctypes.data, a.size, C.byref(h), None))
        return Buffer(h.value)

    @property
    def ptr(self) -> int:
        p = u64()
        check(_L().cc_buffer_device_ptr(self.handle, C.byref(p)))
        return p.value

    @property
    def length(self) -> int:
        p = u64()
        check(_L().cc_buffer_length(self.handle, C.byref(p)))
        return p.value

    @property
    def is_multicast(self) -> bool:
        """a symmetric buffer with an NVLS multicast mapping (cc_comm_symmetric_alloc on an NVSwitch system)"""
        out = C.c_int()
        check(_L().cc_buffer_is_multicast(self.handle, C.byref(out)))
        return bool(out.value)

    def share(self) -> "Buffer":
        """another handle object on the same device buffer (DeviceBuffer.retain, OpenCL.scala:644)"""
        check(_L().cc_buffer_retain(self.handle))
        return Buffer(self.handle)

    def upload(self, host_ptr: int, n_floats: int) -> None:
        """async H2D (host memory must stay alive until the next synchronising call)"""
        ev = u64()
        check(_L().cc_buffer_upload(self.handle, host_ptr, int(n_floats), None, 0, C.byref(ev)))
        check(_L().cc_event_release(ev.value))

    def to_host(self, n_floats: int | None = None, offset: int = 0) -> np.ndarray:
        n = self.length - offset if n_floats is None else int(n_floats)
        out = np.empty(n, dtype=np.float32)
        check(_L().cc_buffer_to_host(self.handle, int(offset), out.ctypes.data, n, None, 0, None))
        return out

    def to_host_async(self, host_ptr: int, n_floats: int, offset: int = 0) -> None:
        ev = u64()
        check(_L().cc_buffer_to_host(self.handle, int(offset), host_ptr, int(n_floats), None, 0, C.byref(ev)))
        check(_L().cc_event_release(ev.value))

    def release(self) -> None:
        h = self.handle
        if h:
            self.handle = 0
            if (_HOT.buffer_release or _hot().buffer_release)(h):
                check(_L().cc_buffer_release(h))  # failed (stale handle): the ctypes call reports the same status with its message

    def __del__(self):
        if self.handle:
            try:
                if _lib._lib is not None:
                    self.release()
            except Exception:
                pass


class _Hot:
    """entry points on per-call hot paths (a launch-bound step is ~3 us, a small read-back ~15: marshalling counts), bound once"""

    __slots__ = ("buffer_release", "do_buffer", "unary", "binary", "tensor_release", "flat_array_into", "flat_buffer", "flat_buffer_release", "native")

    def __init__(self):
        for n in self.__slots__:
            setattr(self, n, None)


_HOT = _Hot()


def _hot() -> _Hot:
    """Handle-returning calls give the handle (> 0) or the negative cc_status; the others give the cc_status.  Bound through the
    _hotcalls CPython extension (csrc/py_hotcalls.c, ~0.1 us per call) when it is built, else through ctypes (>= 0.55 us per call);
    either way the call lands in libcompute_cuda.so."""
    L = _L()
    H = _HOT
    try:
        if os.environ.get("CC_PY_NO_HOTCALLS"):  # A/B switch and test hook: bind everything through ctypes
            raise ImportError("disabled")
        from . import _hotcalls as X

        H.buffer_release, H.do_buffer, H.unary, H.binary, H.tensor_release = X.buffer_release, X.do_buffer, X.unary, X.binary, X.tensor_release
        H.flat_array_into, H.flat_buffer, H.flat_buffer_release = X.flat_array_into, X.flat_buffer, X.flat_buffer_release
        H.native = True
    except ImportError:
        fp = C.POINTER(C.c_float)

        def do_buffer(t: int) -> int:
            h = u64()
            st = L.ct_do_buffer(t, h, None)
            return st if st else h.value

        def unary(op: int, t: int) -> int:
            h = u64()
            st = L.ct_unary(op, t, h)
            return st if st else h.value

        def binary(op: int, l: int, r: int) -> int:
            h = u64()
            st = L.ct_binary(op, l, r, h)
            return st if st else h.value

        def flat_array_into(t: int, out: np.ndarray) -> int:
            return L.ct_flat_array(t, out.ctypes.data, out.size)

        def flat_buffer(t: int):
            p, n = fp(), u64()
            st = L.ct_flat_buffer(t, C.byref(p), C.byref(n))
            if st:
                return st
            addr = C.cast(p, C.c_void_p).value or 0
            view = memoryview((C.c_float * n.value).from_address(addr)).cast("B") if n.value else memoryview(b"")
            return addr, view

        def flat_buffer_release(addr: int) -> int:
            return L.ct_flat_buffer_release(C.cast(addr, fp))

        H.buffer_release, H.do_buffer, H.unary, H.binary, H.tensor_release = L.cc_buffer_release, do_buffer, unary, binary, L.ct_release
        H.flat_array_into, H.flat_buffer, H.flat_buffer_release = flat_array_into, flat_buffer, flat_buffer_release
        H.native = False
    return H


def reduce_sum(src: Buffer, n_floats: int, dst: Buffer) -> None:
    check(_L().cc_reduce_sum(src.handle, int(n_floats), dst.handle, None, 0, None))


def matmul_3xtf32(a: Buffer, b: Buffer, c: Buffer, m: int, n: int, k: int) -> None:
    check(_L().cc_matmul_3xtf32(a.handle, b.handle, c.handle, m, n, k, None, 0, None))


def comm_symmetric_alloc(n_floats: int) -> "Buffer":
    """collective: the same allocation on every rank, mapped into every other rank over NVLink (CUDA IPC)"""
    h = u64()
    check(_L().cc_comm_symmetric_alloc(int(n_floats), C.byref(h)))
    return Buffer(h.value)


def matmul_3xtf32_allgather(a: "Buffer", b: "Buffer", gathered: "Buffer", m_shard: int, n: int, k: int) -> None:
    """collective: row-sharded matmul whose epilogue stores the result blocks into `gathered` on every rank"""
    check(_L().cc_matmul_3xtf32_allgather(a.handle, b.handle, gathered.handle, m_shard, n, k, None, 0, None))


def kernel_cache_limit(max_kernels: int) -> None:
    """0 = unbounded (reference default); otherwise LRU eviction (kernelCacheBuilder.maximumSize, Tensors.scala:1267-1277)"""
    check(_L().cc_kernel_cache_limit(int(max_kernels)))


def kernel_disk_cache(directory: str | None) -> None:
    """opt-in on-disk cubin cache (None / "" = off): a later process skips NVRTC for every structure it has seen before"""
    check(_L().cc_kernel_disk_cache(directory.encode() if directory else None))


def kernel_cache_clear() -> None:
    check(_L().cc_kernel_cache_clear())


def kernel_cache_size() -> int:
    n = u64()
    check(_L().cc_kernel_cache_size(C.byref(n)))
    return n.value


class Graph:
    """A caller-visible sequence of evaluations captured once and replayed with one driver call each time (cc_graph_*):

        with cuda.Graph() as g:
            outs = [expr.doBuffer() for _ in range(50)]     # recorded, not run
        g.launch()                                          # runs the 50 kernels; outs[i] now hold results, refreshed by every launch
    """

    def __init__(self):
        self._h = 0

    def __enter__(self) -> "Graph":
        check(_L().cc_graph_begin())
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        h = u64()
        st = _L().cc_graph_end(C.byref(h))
        if exc_type is None:
            check(st)
            self._h = h.value
        elif st == 0:
            _L().cc_graph_release(h.value)

    def launch(self) -> None:
        check(_L().cc_graph_launch(self._h, None, 0, None))

    @property
    def commands(self) -> int:
        n, b = u64(), u64()
        check(_L().cc_graph_info(self._h, C.byref(n), C.byref(b)))
        return n.value

    def release(self) -> None:
        if self._h:
            h, self._h = self._h, 0
            check(_L().cc_graph_release(h))

    def __del__(self):
        try:
            if _lib._lib is not None:
                self.release()
        except Exception:
            pass


def kernel_cache_lookup(tree_blob: bytes, any_out_shape: bool = False) -> "Kernel | None":
    """probe only (kernelCache.getIfPresent, Tensors.scala:1293): the cached kernel with the blob's structure, or None"""
    h = u64()
    check(_L().cc_kernel_cache_lookup(tree_blob, len(tree_blob), 1 if any_out_shape else 0, C.byref(h)))
    return Kernel(h.value) if h.value else None


def compile_blob(tree_blob: bytes) -> "Kernel":
    """cc_compile on a raw tree blob (what a JVM front end hands over)"""
    h = u64()
    check(_L().cc_compile(tree_blob, len(tree_blob), C.byref(h)))
    return Kernel(h.value)


def shard_rows(rows: int, n_ranks: int, rank: int) -> tuple[int, int]:
    first, count = C.c_int64(), C.c_int64()
    check(_L().cc_shard_rows(rows, n_ranks, rank, C.byref(first), C.byref(count)))
    return first.value, count.value


def shard_agree(value: int) -> bool:
    """collective: did every rank pass the same value?"""
    out = C.c_int()
    check(_L().cc_shard_agree(value, C.byref(out)))
    return bool(out.value)


def comm_info() -> tuple[int, int]:
    n, r = C.c_int(), C.c_int()
    check(_L().cc_comm_info(C.byref(n), C.byref(r)))
    return n.value, r.value


def set_operand_cache(on: bool) -> None:
    """keep (True, default) or drop (False) the TF32 hi / lo panels of recently used, unchanged B operands"""
    check(_L().cc_set_operand_cache(1 if on else 0))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_L().cc_comm_unique_id(buf))
    return buf.raw


def comm_init(unique_id: bytes, n_ranks: int, rank: int) -> None:
    check(_L().cc_comm_init(C.create_string_buffer(unique_id, 128), int(n_ranks), int(rank)))


def comm_enable_peer() -> None:
    """map the NVLink peer mailboxes; afterwards small all-reduces and the sharded Tensor.sum bypass NCCL"""
    check(_L().cc_comm_enable_peer())


def comm_route_peer(on: bool) -> None:
    check(_L().cc_comm_route_peer(1 if on else 0))


def comm_peer_enabled() -> bool:
    v = C.c_int()
    check(_L().cc_comm_peer_enabled(C.byref(v)))
    return bool(v.value)


def reduce_sum_allreduce(src: Buffer, n_floats: int, dst: Buffer) -> None:
    check(_L().cc_reduce_sum_allreduce(src.handle, int(n_floats), dst.handle, None, 0, None))


def comm_destroy() -> None:
    check(_L().cc_comm_destroy())


def allreduce_sum(buf: Buffer, n_floats: int) -> None:
    check(_L().cc_allreduce_sum(buf.handle, int(n_floats), None, 0, None))


def allgather(send: Buffer, recv: Buffer, n_floats_per_rank: int) -> None:
    check(_L().cc_allgather(send.handle, recv.handle, int(n_floats_per_rank), None, 0, None))


def broadcast(buf: Buffer, n_floats: int, root: int = 0) -> None:
    check(_L().cc_broadcast(buf.handle, int(n_floats), int(root), None, 0, None))


# ---- Tensor (Tensors.scala:395-1261) ---------------------------------------------------------------------------------------


def _flatten(elements):
    """TensorBuilder (Tensors.scala:44-94): nested sequences -> (shape, row-major flat floats); ragged input is an
    IllegalArgumentException (TensorsSpec.scala:67-73)."""
    if isinstance(elements, np.ndarray):
        return tuple(elements.shape), np.ascontiguousarray(elements, dtype=np.float32).reshape(-1)
    if isinstance(elements, (int, float, np.floating, np.integer)):
        return (), np.asarray([elements], dtype=np.float32)
    subs = [_flatten(e) for e in elements]
    if not subs:
        return (0,), np.zeros(0, dtype=np.float32)
    for s, _ in subs:
        if s != subs[0][0]:
            raise IllegalArgumentException(-1, "tensor literal is not rectangular")
    return (len(subs),) + subs[0][0], np.concatenate([f for _, f in subs]).astype(np.float32)


class Tensor:
    """`cuda.Tensor` — lazily evaluated N-dimensional float32 array."""

    __slots__ = ("_h", "_shape", "__weakref__")

    def __init__(self, elements=None, padding: float = 0.0, *, _handle: int | None = None):
        self._shape = None
        if _handle is not None:
            self._h = _handle
            return
        shape, flat = _flatten(elements)
        sa, rank = _shape_arg(shape)
        h = u64()
        check(_L().ct_from_host(flat.ctypes.data, sa, rank, float(padding), C.byref(h)))
        self._h = h.value

    # -- construction (object Tensor) --
    @staticmethod
    def _wrap(h: u64) -> "Tensor":
        return Tensor._of(h.value)

    @staticmethod
    def _of(handle: int) -> "Tensor":
        t = object.__new__(Tensor)
        t._h = handle
        t._shape = None
        return t

    @staticmethod
    def scalar(value: float, padding: float = 0.0) -> "Tensor":
        h = u64()
        check(_L().ct_scalar(float(value), float(padding), C.byref(h)))
        return Tensor._wrap(h)

    @staticmethod
    def fill(value: float, shape: Sequence[int], padding: float = 0.0) -> "Tensor":
        sa, rank = _shape_arg(shape)
        h = u64()
        check(_L().ct_fill(float(value), sa, rank, float(padding), C.byref(h)))
        return Tensor._wrap(h)

    @staticmethod
    def random(shape: Sequence[int], seed: int, padding: float = 0.0) -> "Tensor":
        sa, rank = _shape_arg(shape)
        h = u64()
        check(_L().ct_random(sa, rank, np.int32(np.uint32(seed & 0xFFFFFFFF)).item(), float(padding), C.byref(h)))
        return Tensor._wrap(h)

    @staticmethod
    def randomNormal(shape: Sequence[int], seed: int, padding: float = 0.0) -> "Tensor":
        sa, rank = _shape_arg(shape)
        h = u64()
        check(_L().ct_random_normal(sa, rank, np.int32(np.uint32(seed & 0xFFFFFFFF)).item(), float(padding), C.byref(h)))
        return Tensor._wrap(h)

    @staticmethod
    def fromBuffer(buf: Buffer, shape: Sequence[int], padding: float = 0.0) -> "Tensor":
        sa, rank = _shape_arg(shape)
        h = u64()
        check(_L().ct_from_buffer(buf.handle, sa, rank, float(padding), C.byref(h)))
        return Tensor._wrap(h)

    @staticmethod
    def _un(op: str, t: "Tensor") -> "Tensor":
        h = (_HOT.unary or _hot().unary)(_UNARY[op], t._h)
        if h < 0:
            check(h)
        return Tensor._of(h)

    @staticmethod
    def _bin(op: str, l: "Tensor", r: "Tensor") -> "Tensor":
        h = (_HOT.binary or _hot().binary)(_BINARY[op], l._h, r._h)
        if h < 0:
            check(h)
        return Tensor._of(h)

    abs = staticmethod(lambda t: _unop(12, t))
    sqrt = staticmethod(lambda t: _unop(14, t))
    tanh = staticmethod(lambda t: _unop(13, t))
    exp = staticmethod(lambda t: _unop(10, t))
    log = staticmethod(lambda t: _unop(11, t))
    min = staticmethod(lambda l, r: _binop(20, l, r))
    max = staticmethod(lambda l, r: _binop(21, l, r))

    @staticmethod
    def join(tensors: Iterable["Tensor"], dimension: int | None = None) -> "Tensor":
        ts = list(tensors)
        arr = (u64 * max(1, len(ts)))(*[t._h for t in ts])
        h = u64()
        if dimension is None:
            check(_L().ct_join(arr, len(ts), C.byref(h)))
        else:
            check(_L().ct_join_dim(arr, len(ts), int(dimension), C.byref(h)))
        return Tensor._wrap(h)

    # -- operators --
    def __add__(self, o):
        return _binop(22, self, o)

    def __sub__(self, o):
        return _binop(23, self, o)

    def __mul__(self, o):
        return _binop(24, self, o)

    def __truediv__(self, o):
        return _binop(25, self, o)

    def __mod__(self, o):
        return _binop(26, self, o)

    def __neg__(self):
        return _unop(15, self)

    def __pos__(self):
        return self

    # -- delayed --
    def _shape_op(self, fn, shape) -> "Tensor":
        sa, rank = _shape_arg(shape)
        h = u64()
        check(fn(self._h, sa, rank, C.byref(h)))
        return Tensor._wrap(h)

    def broadcast(self, newShape) -> "Tensor":
        return self._shape_op(_L().ct_broadcast, newShape)

    def reshape(self, newShape) -> "Tensor":
        return self._shape_op(_L().ct_reshape, newShape)

    def scale(self, newShape) -> "Tensor":
        return self._shape_op(_L().ct_scale, newShape)

    def translate(self, offset: Sequence[float], newShape: Sequence[int] | None = None) -> "Tensor":
        off = (C.c_double * max(1, len(offset)))(*[float(o) for o in offset])
        h = u64()
        if newShape is None:
            check(_L().ct_translate(self._h, off, len(offset), None, -1, C.byref(h)))
        else:
            sa, rank = _shape_arg(newShape)
            check(_L().ct_translate(self._h, off, len(offset), sa, rank, C.byref(h)))
        return Tensor._wrap(h)

    def permute(self, dimensions: Sequence[int]) -> "Tensor":
        sa, n = _shape_arg(dimensions)
        h = u64()
        check(_L().ct_permute(self._h, sa, n, C.byref(h)))
        return Tensor._wrap(h)

    def transpose(self) -> "Tensor":
        h = u64()
        check(_L().ct_transpose(self._h, C.byref(h)))
        return Tensor._wrap(h)

    def split(self, dimension: int) -> list["Tensor"]:
        n = C.c_int()
        check(_L().ct_split(self._h, int(dimension), None, 0, C.byref(n)))
        arr = (u64 * max(1, n.value))()
        check(_L().ct_split(self._h, int(dimension), arr, n.value, C.byref(n)))
        return [Tensor(_handle=arr[i]) for i in range(n.value)]

    def sum(self) -> "Tensor":
        h = u64()
        check(_L().ct_sum(self._h, C.byref(h)))
        return Tensor._wrap(h)

    def reduce(self, monoid: str) -> "Tensor":
        """reduce(MonoidPrograms) (Tensors.scala:308-311, 673-766): fold the whole tensor with "+", "*", "min" or "max";
        an inline operand is fused into the fold kernel"""
        if monoid not in ("+", "*", "min", "max"):
            raise IllegalArgumentException(-1, f"not a monoid: {monoid!r}")
        h = u64()
        check(_L().ct_reduce(self._h, _BINARY[monoid], C.byref(h)))
        return Tensor._wrap(h)

    def product(self) -> "Tensor":
        return self.reduce("*")

    def nonInline(self) -> "Tensor":
        h = u64()
        check(_L().ct_non_inline(self._h, C.byref(h)))
        return Tensor._wrap(h)

    def doCache(self) -> "Tensor":
        """evaluates now and pins the buffer until the returned tensor is released (Tensors.scala:642-666)"""
        h = u64()
        check(_L().ct_do_cache(self._h, C.byref(h)))
        return Tensor._wrap(h)

    # -- properties --
    @property
    def shape(self) -> tuple:
        s = self._shape
        if s is None:  # a tensor's shape never changes: asked of the library once
            r = C.c_int()
            check(_L().ct_rank(self._h, C.byref(r)))
            arr = (i32 * max(1, r.value))()
            check(_L().ct_shape(self._h, arr, r.value))
            s = self._shape = tuple(arr[i] for i in range(r.value))
        return s

    @property
    def padding(self) -> float:
        p = C.c_float()
        check(_L().ct_padding(self._h, C.byref(p)))
        return p.value

    # -- slow actions --
    def flatArray(self) -> np.ndarray:
        out = np.empty(math.prod(self.shape), dtype=np.float32)
        st = (_HOT.flat_array_into or _hot().flat_array_into)(self._h, out)
        if st:
            check(st)
        return out

    def flatBuffer(self) -> HostBuffer:
        """evaluate and read back into pooled pinned memory (no pageable staging copy): `with t.flatBuffer() as a: ...`"""
        r = (_HOT.flat_buffer or _hot().flat_buffer)(self._h)
        if r.__class__ is int:
            check(r)
        return HostBuffer(r[0], r[1])

    def flatArrayInto(self, host_ptr: int, capacity_floats: int) -> None:
        check(_L().ct_flat_array(self._h, host_ptr, int(capacity_floats)))

    def toString(self) -> str:
        need = u64()
        check(_L().ct_to_string(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(_L().ct_to_string(self._h, buf, need.value, None))
        return buf.value.decode()

    __str__ = toString

    def doBuffer(self) -> Buffer:
        h = (_HOT.do_buffer or _hot().do_buffer)(self._h)
        if h < 0:
            check(h)
        return Buffer(h)

    def compile(self) -> Kernel:
        h = u64()
        check(_L().ct_compile(self._h, C.byref(h)))
        return Kernel(h.value)

    def treeBlob(self) -> bytes:
        """the tree blob compile() hands to cc_compile_ex (definitions attached) — what CudaTreeWriter.scala must write"""
        need = u64()
        check(_L().ct_tree_blob(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(_L().ct_tree_blob(self._h, buf, need.value, C.byref(need)))
        return buf.raw[: need.value]

    # ---- sharding over the GPUs of one box (include/compute_cuda.h: ct_shard / ct_gather) ----

    def shard(self) -> "Tensor":
        """declares this tensor to be THIS rank's row block of a tensor sharded along its leading axis"""
        h = u64()
        check(_L().ct_shard(self._h, C.byref(h)))
        return Tensor._wrap(h)

    @property
    def distribution(self) -> str:
        d = C.c_int()
        check(_L().ct_distribution(self._h, C.byref(d)))
        return ("whole", "row block", "partial sum")[d.value]

    def gather(self, zero_copy: bool = False) -> "Tensor":
        """row block -> the whole tensor on every rank (a sharded matmul result is gathered by the contraction's own epilogue)"""
        h = u64()
        check(_L().ct_gather(self._h, 1 if zero_copy else 0, C.byref(h)))
        return Tensor._wrap(h)

    def release(self) -> None:
        try:
            h = self._h
        except AttributeError:  # construction failed before the handle existed
            return
        if h:
            self._h = 0
            st = (_HOT.tensor_release or _hot().tensor_release)(h)
            if st:
                check(st)

    def __del__(self):
        try:
            if _lib._lib is not None:
                self.release()
        except Exception:
            pass


def _unop(code: int, t: Tensor) -> Tensor:
    """one elementwise node (codes = _UNARY's: the C ABI's CT_* constants); the operators above call this directly"""
    h = (_HOT.unary or _hot().unary)(code, t._h)
    if h < 0:
        check(h)
    r = object.__new__(Tensor)
    r._h = h
    r._shape = None
    return r


def _binop(code: int, l: Tensor, r: Tensor) -> Tensor:
    h = (_HOT.binary or _hot().binary)(code, l._h, r._h)
    if h < 0:
        check(h)
    t = object.__new__(Tensor)
    t._h = h
    t._shape = None
    return t


assert (_UNARY["exp"], _UNARY["log"], _UNARY["abs"], _UNARY["tanh"], _UNARY["sqrt"], _UNARY["neg"]) == (10, 11, 12, 13, 14, 15)
assert [_BINARY[k] for k in ("min", "max", "+", "-", "*", "/", "%")] == [20, 21, 22, 23, 24, 25, 26]


def live_tensors() -> int:
    n = C.c_int64()
    check(_L().ct_live_tensors(C.byref(n)))
    return n.value
