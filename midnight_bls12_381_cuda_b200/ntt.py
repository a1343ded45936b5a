"""GpuNttContext -- mirrors core/ntt.rs:303-1463 over the C ABI.

The reference's registered backend ignores ordering / coset_gen / stream; this one
honours them, so forward_coset_ntt really evaluates on the coset.  The `*_auto` CPU/GPU dispatchers
(core/ntt.rs:1879-1990) are intentionally absent: north_star removes the hybrid threshold.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L
from .stream import DeviceVec, ManagedStream, ensure_backend_loaded, set_device
from .types import TypeConverter

FORWARD, INVERSE = 0, 1
kNN, kNR, kRN, kRR, kNM, kMN = range(6)

# 7^((r-1)/2^32) in STANDARD form: what icicle's get_root_of_unity yields upstream (core/ntt.rs:412-413)
_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_ROOT_2_32 = pow(7, (_R - 1) >> 32, _R)


class NttError(RuntimeError):
    """core/ntt.rs:92-150"""


def get_root_of_unity(size: int) -> np.ndarray:
    """standard-form primitive `size`-th root (size = 2^k, k <= 32) as 4 LE limbs."""
    k = size.bit_length() - 1
    if size <= 0 or (1 << k) != size or k > 32:
        raise NttError(f"NTT size must be a power of 2 <= 2^32, got {size}")
    w = pow(_ROOT_2_32, 1 << (32 - k), _R)
    return np.array([(w >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class NttHandle:
    """core/ntt.rs:1409-1463"""

    def __init__(self, stream, out, keep):
        self._stream, self._out, self._keep = stream, out, keep

    def size(self) -> int:
        return self._out.shape[0]

    def wait(self) -> np.ndarray:
        self._stream.synchronize()
        self._stream.destroy()
        self._keep = None
        return self._out


_domain_lock = threading.Lock()
_domain_log = 0            # device 0 (kept as a plain int: tests reset it)
_domain_logs: dict[int, int] = {}   # other devices: the backend keeps one domain per device


class GpuNttContext:
    def __init__(self, max_log_size: int, device_id: int = 0, ordering: int = kNN):
        try:
            ensure_backend_loaded()
            set_device(device_id)
        except Exception as e:  # noqa: BLE001
            raise NttError(f"backend init failed: {e}") from e
        self._max_log = max_log_size
        self.device_id = device_id
        self.ordering = ordering
        self._ensure_domain_initialized(max_log_size, device_id)

    @staticmethod
    def _ensure_domain_initialized(k: int, device_id: int = 0) -> None:
        """core/ntt.rs:380-442: one domain per device (the current device's), grown on demand."""
        global _domain_log
        with _domain_lock:
            have = _domain_log if device_id == 0 else _domain_logs.get(device_id, 0)
            if have >= k:
                return
            lib = L.lib()
            if have:
                lib.b381_ntt_release_domain()
            root = get_root_of_unity(1 << k)
            cfg = L.NTTInitDomainConfig()
            code = lib.b381_ntt_init_domain(L.ptr(root), C.byref(cfg))
            if code != 0:
                raise NttError(f"init_domain(2^{k}): {L.ERROR_NAMES.get(code, code)}")
            if device_id == 0:
                _domain_log = k
            else:
                _domain_logs[device_id] = k

    def max_log_size(self) -> int:
        return self._max_log

    # -- core call -----------------------------------------------------------
    def _cfg(self, batch=1, coset_gen=None, on_device=False, stream=None, is_async=False, ordering=None, columns=False):
        cfg = L.lib().b381_default_ntt_config()
        cfg.batch_size = batch
        cfg.columns_batch = columns
        cfg.ordering = self.ordering if ordering is None else ordering
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = on_device
        cfg.is_async = is_async
        if stream is not None:
            cfg.stream = stream.handle
        if coset_gen is not None:
            g = np.asarray(coset_gen, dtype=np.uint64).reshape(4)
            for i in range(4):
                cfg.coset_gen.l[i] = int(g[i])
        return cfg

    def _run(self, src, dst, size, direction, cfg):
        if size == 0:
            return
        if size & (size - 1):
            raise NttError(f"NTT size must be power of 2, got {size}")
        if size.bit_length() - 1 > self._max_log:
            raise NttError(f"size 2^{size.bit_length() - 1} exceeds domain 2^{self._max_log}")
        code = L.lib().b381_ntt(L.ptr(src), size, direction, C.byref(cfg), L.ptr(dst))
        if code != 0:
            raise NttError(f"ntt: {L.ERROR_NAMES.get(code, code)}")

    def _host(self, data, direction, batch=1, poly_size=None, coset_gen=None, inplace=False):
        a = TypeConverter.scalar_slice_as_icicle(data)
        size = a.shape[0] // batch if poly_size is None else poly_size
        if poly_size is not None and a.shape[0] % poly_size:
            raise NttError("batch length is not a multiple of poly_size")
        batch = a.shape[0] // size if size else batch
        out = a if inplace else np.empty_like(a)
        self._run(a, out, size, direction, self._cfg(batch=batch, coset_gen=coset_gen))
        return out

    # -- host-slice API (names of core/ntt.rs:453-1390) ----------------------
    def forward_ntt(self, coefficients):
        return self._host(coefficients, FORWARD)

    def inverse_ntt(self, evaluations):
        return self._host(evaluations, INVERSE)

    def forward_ntt_inplace(self, data):
        self._host(data, FORWARD, inplace=True)

    def inverse_ntt_inplace(self, data):
        self._host(data, INVERSE, inplace=True)

    def forward_ntt_batch(self, batch, poly_size):
        return self._host(batch, FORWARD, poly_size=poly_size)

    def inverse_ntt_batch(self, batch, poly_size):
        return self._host(batch, INVERSE, poly_size=poly_size)

    def forward_ntt_batch_inplace(self, batch, poly_size):
        self._host(batch, FORWARD, poly_size=poly_size, inplace=True)

    def inverse_ntt_batch_inplace(self, batch, poly_size):
        self._host(batch, INVERSE, poly_size=poly_size, inplace=True)

    def forward_coset_ntt(self, coefficients, coset_gen):
        return self._host(coefficients, FORWARD, coset_gen=coset_gen)

    def inverse_coset_ntt(self, evaluations, coset_gen):
        return self._host(evaluations, INVERSE, coset_gen=coset_gen)

    def forward_coset_ntt_inplace(self, data, coset_gen):
        self._host(data, FORWARD, coset_gen=coset_gen, inplace=True)

    def inverse_coset_ntt_inplace(self, data, coset_gen):
        self._host(data, INVERSE, coset_gen=coset_gen, inplace=True)

    def forward_coset_ntt_batch(self, batch, poly_size, coset_gen):
        return self._host(batch, FORWARD, poly_size=poly_size, coset_gen=coset_gen)

    def inverse_coset_ntt_batch(self, batch, poly_size, coset_gen):
        return self._host(batch, INVERSE, poly_size=poly_size, coset_gen=coset_gen)

    def forward_coset_ntt_batch_inplace(self, batch, poly_size, coset_gen):
        self._host(batch, FORWARD, poly_size=poly_size, coset_gen=coset_gen, inplace=True)

    def inverse_coset_ntt_batch_inplace(self, batch, poly_size, coset_gen):
        self._host(batch, INVERSE, poly_size=poly_size, coset_gen=coset_gen, inplace=True)

    # -- device-resident API (core/ntt.rs:610-919) ---------------------------
    def ntt_on_device(self, device_data, direction, size=None, batch=1, coset_gen=None, stream=None, is_async=False, ordering=None):
        """in place on a DeviceVec / torch tensor / raw pointer holding batch*size Fr elements."""
        if size is None:
            size = len(device_data) // batch
        self._run(device_data, device_data, size, direction,
                  self._cfg(batch=batch, coset_gen=coset_gen, on_device=True, stream=stream, is_async=is_async, ordering=ordering))

    def forward_ntt_on_device(self, device_data):
        self.ntt_on_device(device_data, FORWARD)

    def inverse_ntt_on_device(self, device_data):
        self.ntt_on_device(device_data, INVERSE)

    def ntt_batch_on_device(self, device_data, poly_size, direction):
        self.ntt_on_device(device_data, direction, size=poly_size, batch=len(device_data) // poly_size)

    def forward_ntt_batch_on_device(self, device_data, poly_size):
        self.ntt_batch_on_device(device_data, poly_size, FORWARD)

    def inverse_ntt_batch_on_device(self, device_data, poly_size):
        self.ntt_batch_on_device(device_data, poly_size, INVERSE)

    def coset_ntt_on_device(self, device_data, direction, coset_gen):
        self.ntt_on_device(device_data, direction, coset_gen=coset_gen)

    def coset_ntt_batch_on_device(self, device_data, poly_size, direction, coset_gen):
        self.ntt_on_device(device_data, direction, size=poly_size, batch=len(device_data) // poly_size, coset_gen=coset_gen)

    def ntt_on_device_async(self, device_data, direction, stream: ManagedStream):
        self.ntt_on_device(device_data, direction, stream=stream, is_async=True)

    def ntt_batch_on_device_async(self, device_data, poly_size, direction, stream: ManagedStream):
        self.ntt_on_device(device_data, direction, size=poly_size, batch=len(device_data) // poly_size, stream=stream, is_async=True)

    # -- async host API (core/ntt.rs:945-1040) -------------------------------
    def _async(self, data, direction) -> NttHandle:
        a = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(data))
        st = ManagedStream.create()
        d = DeviceVec(a.shape[0], 32)
        L.check(L.lib().b381_copy_to_device_async(L.ptr(d), L.ptr(a), a.nbytes, st.handle), "h2d")
        self.ntt_on_device(d, direction, size=a.shape[0], stream=st, is_async=True)
        out = np.empty_like(a)
        L.check(L.lib().b381_copy_to_host_async(L.ptr(out), L.ptr(d), a.nbytes, st.handle), "d2h")
        return NttHandle(st, out, (a, d))

    def forward_ntt_async(self, coefficients) -> NttHandle:
        return self._async(coefficients, FORWARD)

    def inverse_ntt_async(self, evaluations) -> NttHandle:
        return self._async(evaluations, INVERSE)
