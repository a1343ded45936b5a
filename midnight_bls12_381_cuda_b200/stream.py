"""ManagedStream / DeviceVec -- mirrors core/stream.rs:96-198 (RAII CUDA stream) and the slice of
icicle_runtime::memory::DeviceVec the reference's core/ uses (device_malloc, copy_from_host,
copy_to_host, len)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class GpuError(L.B381Error):
    """core/backend.rs:36"""


def ensure_backend_loaded() -> None:
    """core/backend.rs:75-97.  Loads the CUDA library; raises if it is missing -- there is no CPU path."""
    L.lib()


def is_gpu_available() -> bool:
    n = C.c_int(0)
    try:
        return L.lib().b381_device_count(C.byref(n)) == 0 and n.value > 0
    except L.BackendNotBuilt:
        return False


def set_device(device_id: int = 0) -> None:
    L.check(L.lib().b381_set_device(device_id), "set_device")


class ManagedStream:
    def __init__(self, handle: int | None, owned: bool):
        self._h = handle
        self._owned = owned
        self._destroyed = False

    @classmethod
    def create(cls) -> "ManagedStream":
        h = C.c_void_p()
        L.check(L.lib().b381_stream_create(C.byref(h)), "stream create")
        return cls(h.value, True)

    @classmethod
    def default_stream(cls) -> "ManagedStream":
        return cls(None, False)

    @property
    def handle(self) -> C.c_void_p:
        return C.c_void_p(self._h)

    def synchronize(self) -> None:
        L.check(L.lib().b381_stream_synchronize(self.handle), "stream synchronize")

    def destroy(self) -> None:
        if self._owned and not self._destroyed:
            L.check(L.lib().b381_stream_destroy(self.handle), "stream destroy")
        self._destroyed = True

    def is_destroyed(self) -> bool:
        return self._destroyed

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class DeviceVec:
    """Owned device buffer of `count` elements of `elem_bytes` bytes."""

    def __init__(self, count: int, elem_bytes: int):
        self.count, self.elem_bytes, self._owned = count, elem_bytes, True
        p = C.c_void_p()
        L.check(L.lib().b381_malloc(C.byref(p), max(1, count * elem_bytes)), "device_malloc")
        self.ptr = p.value

    @classmethod
    def borrow(cls, ptr: int, count: int, elem_bytes: int) -> "DeviceVec":
        """non-owning view of device memory somebody else allocated (a torch tensor, an ICICLE DeviceSlice)"""
        v = cls.__new__(cls)
        v.count, v.elem_bytes, v.ptr, v._owned = count, elem_bytes, ptr, False
        return v

    @classmethod
    def from_host(cls, arr: np.ndarray, elem_bytes: int) -> "DeviceVec":
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes % elem_bytes == 0
        v = cls(arr.nbytes // elem_bytes, elem_bytes)
        v.copy_from_host(arr)
        return v

    def __len__(self) -> int:
        return self.count

    @property
    def nbytes(self) -> int:
        return self.count * self.elem_bytes

    def data_ptr(self) -> int:
        return self.ptr

    def copy_from_host(self, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes == self.nbytes
        L.check(L.lib().b381_copy_to_device(C.c_void_p(self.ptr), L.ptr(arr), arr.nbytes), "copy_from_host")

    def copy_to_host(self, dtype=np.uint64) -> np.ndarray:
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        L.check(L.lib().b381_copy_to_host(L.ptr(out), C.c_void_p(self.ptr), self.nbytes), "copy_to_host")
        return out

    def free(self) -> None:
        if self.ptr and self._owned:
            L.lib().b381_free(C.c_void_p(self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """uint64 result buffer in page-locked host memory.  An is_async call copies its result device -> host on the
    call's stream; into pageable memory that copy blocks the calling thread until the whole call has run, into pinned
    memory it does not -- which is what lets two handles overlap (core/msm.rs:715-798 keeps its results in DeviceVecs
    for the same reason).  Blocks are recycled through a free list: cudaHostAlloc / cudaFreeHost cost a device-wide
    synchronisation each, which would serialise the very calls this exists for."""

    _free: dict = {}          # words -> [pointer, ...]

    def __init__(self, shape):
        self.shape = tuple(np.atleast_1d(shape))
        n = max(1, int(np.prod(self.shape)))
        self._words = n
        cached = PinnedArray._free.get(n)
        if cached:
            self._p = cached.pop()
        else:
            p = C.c_void_p()
            L.check(L.lib().b381_host_alloc_pinned(C.byref(p), 8 * n), "host_alloc_pinned")
            self._p = p.value
        self.array = np.ctypeslib.as_array(C.cast(C.c_void_p(self._p), C.POINTER(C.c_uint64)), shape=(n,)).reshape(self.shape)
        self.array[...] = 0

    def take(self) -> np.ndarray:
        """pageable copy of the contents; the pinned block goes back to the free list"""
        out = self.array.copy()
        self.free()
        return out

    def free(self) -> None:
        if self._p:
            self.array = None
            PinnedArray._free.setdefault(self._words, []).append(self._p)
            self._p = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
