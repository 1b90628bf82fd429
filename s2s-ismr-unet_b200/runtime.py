"""Thin host-side runtime over the C ABI: streams, device / pinned buffers, timed regions.

Everything here is plumbing around include/s2s_unet.h; no arithmetic happens on the host.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import call


def device_count() -> int:
    n = C.c_int(0)
    try:
        call("s2s_device_count", C.byref(n))
    except _lib.S2SError:
        return 0
    return n.value


def set_device(dev: int) -> None:
    call("s2s_set_device", int(dev))


def current_device() -> int:
    """The calling THREAD's current CUDA device (cudaSetDevice is per-thread state)."""
    d = C.c_int(0)
    call("s2s_get_device", C.byref(d))
    return d.value


class Stream:
    def __init__(self):
        p = C.c_void_p()
        call("s2s_stream_create", C.byref(p))
        self.ptr = p.value
        self._fin = weakref.finalize(self, _lib.load().s2s_stream_destroy, C.c_void_p(self.ptr))

    def synchronize(self) -> None:
        call("s2s_stream_sync", C.c_void_p(self.ptr))


class Event:
    def __init__(self):
        p = C.c_void_p()
        call("s2s_event_create", C.byref(p))
        self.ptr = p.value
        self._fin = weakref.finalize(self, _lib.load().s2s_event_destroy, C.c_void_p(self.ptr))

    def record(self, stream: Stream) -> None:
        call("s2s_event_record", C.c_void_p(self.ptr), C.c_void_p(stream.ptr))

    def elapsed_ms(self, end: "Event") -> float:
        ms = C.c_float(0)
        call("s2s_event_elapsed_ms", C.c_void_p(self.ptr), C.c_void_p(end.ptr), C.byref(ms))
        return ms.value


class DeviceBuffer:
    """Owning device allocation of `nbytes`; `ptr` is the raw device address."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        call("s2s_dev_alloc", C.byref(p), C.c_size_t(max(self.nbytes, 1)))
        self.ptr = p.value
        self._stream = None      # the stream this buffer was last used on through this object
        self._fin = weakref.finalize(self, _lib.load().s2s_dev_free, C.c_void_p(self.ptr))

    def free(self) -> None:
        """Return the block to the library's cache.  s2s_dev_free does not synchronise (a device-wide synchronisation would
        break stream captures of other trial threads): the stream the buffer was last used on is synchronised here; work
        enqueued on any other stream must have been synchronised by the caller."""
        if self._fin.alive and self._stream is not None:
            try:
                self._stream.synchronize()
            except Exception:
                pass
        self._fin()

    def upload(self, arr: np.ndarray, stream: Stream, offset: int = 0) -> None:
        arr = np.ascontiguousarray(arr)
        assert offset + arr.nbytes <= self.nbytes, "upload overflows the device buffer"
        self._stream = stream
        call("s2s_memcpy_h2d", C.c_void_p(self.ptr + offset), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes),
             C.c_void_p(stream.ptr))

    def download(self, shape, dtype, stream: Stream, offset: int = 0) -> np.ndarray:
        out = np.empty(shape, dtype)
        assert offset + out.nbytes <= self.nbytes, "download overruns the device buffer"
        self._stream = stream
        call("s2s_memcpy_d2h", out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr + offset), C.c_size_t(out.nbytes),
             C.c_void_p(stream.ptr))
        stream.synchronize()
        return out

    @classmethod
    def from_array(cls, arr: np.ndarray, stream: Stream) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        b = cls(arr.nbytes)
        b.upload(arr, stream)
        stream.synchronize()     # the host array may be pageable and go away
        return b


_PINNED_RANGES: dict[int, int] = {}      # base address -> nbytes of every live pinned allocation
_PINNED_HITS: set = set()                # (address, nbytes) of views already found inside a live pinned allocation


def is_pinned(arr: np.ndarray) -> bool:
    """True if `arr` is a C-contiguous view into memory allocated by PinnedArray (direct async H2D source)."""
    if not isinstance(arr, np.ndarray) or not arr.flags["C_CONTIGUOUS"]:
        return False
    p, n = arr.ctypes.data, arr.nbytes
    if (p, n) in _PINNED_HITS:                    # the training loop passes the same pinned batches again and again
        return True
    ok = any(b <= p and p + n <= b + sz for b, sz in _PINNED_RANGES.items())
    if ok and len(_PINNED_HITS) < 4096:
        _PINNED_HITS.add((p, n))
    return ok


def _unpin(ptr, free_fn):
    _PINNED_RANGES.pop(ptr, None)
    _PINNED_HITS.clear()
    free_fn(C.c_void_p(ptr))


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory; batches built here are uploaded without a staging copy."""
    pa = PinnedArray(shape, dtype)
    arr = pa.array
    _KEEP.append(pa)
    return arr


_KEEP: list = []


class PinnedArray:
    """Page-locked host array (numpy view) for asynchronous H2D / D2H copies."""

    def __init__(self, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        call("s2s_host_alloc", C.byref(p), C.c_size_t(max(n, 1)))
        self.ptr = p.value
        buf = (C.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)
        _PINNED_RANGES[self.ptr] = max(n, 1)
        self._fin = weakref.finalize(self, _unpin, self.ptr, _lib.load().s2s_host_free)


def h2d(dst_ptr: int, host_ptr: int, nbytes: int, stream: Stream) -> None:
    call("s2s_memcpy_h2d", C.c_void_p(dst_ptr), C.c_void_p(host_ptr), C.c_size_t(nbytes), C.c_void_p(stream.ptr))


def d2h(host_ptr: int, src_ptr: int, nbytes: int, stream: Stream) -> None:
    call("s2s_memcpy_d2h", C.c_void_p(host_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes), C.c_void_p(stream.ptr))


def d2d(dst_ptr: int, src_ptr: int, nbytes: int, stream: Stream) -> None:
    call("s2s_memcpy_d2d", C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes), C.c_void_p(stream.ptr))
