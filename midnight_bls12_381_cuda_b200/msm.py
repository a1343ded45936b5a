"""GpuMsmContext and friends -- mirrors core/msm.rs:92-1631 over the C ABI (include/b381.h).

Differences from the reference, all deliberate (SURVEY.md 8a "defects not to copy"):
  * batch_size and precompute_factor are honoured by the backend (the reference ignores both);
  * no CPU fallback / size threshold (north_star) -- every call runs on the GPU or raises MsmError.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib as L
from .stream import DeviceVec, ManagedStream, PinnedArray, ensure_backend_loaded, set_device
from .types import G1_AFFINE_BYTES, G1_PROJECTIVE_BYTES, G2_AFFINE_BYTES, G2_PROJECTIVE_BYTES, TypeConverter


class MsmError(RuntimeError):
    """core/msm.rs:92-127"""


class PrecomputedBases:
    """core/msm.rs:174-262: device-resident bases, optionally expanded by precompute_factor.

    Precomputed buffers hold point i's multiples interleaved at [i*factor + k] (upstream ICICLE's layout), so an MSM
    over fewer scalars than original_size uses a prefix of the same buffer, as core/msm.rs:654-661 expects.  `window`
    is the cfg.c the table was built with (0 = the backend's fixed default for precomputed bases); every MSM over these
    bases is issued with the same value, whatever the context's own window is."""

    def __init__(self, buffer: DeviceVec, size: int, factor: int = 1, window: int = 0):
        self.buffer, self._size, self._factor, self.window = buffer, size, factor, window

    def is_precomputed(self) -> bool:
        return self._factor > 1

    def factor(self) -> int:
        return self._factor

    def original_size(self) -> int:
        return self._size

    def buffer_size(self) -> int:
        return len(self.buffer)

    def __len__(self) -> int:
        return self._size

    def required_size_for_scalars(self, num_scalars: int) -> int:
        if num_scalars > self._size:
            raise MsmError(f"{num_scalars} scalars > {self._size} bases")
        return num_scalars * self._factor


class _Handle:
    def __init__(self, stream: ManagedStream, result: np.ndarray, keep):
        self._stream, self._result, self._keep = stream, result, keep

    def _finish(self):
        self._stream.synchronize()
        self._stream.destroy()
        self._keep = None
        if isinstance(self._result, PinnedArray):
            self._result = self._result.take()
        return self._result


class MsmHandle(_Handle):
    """core/msm.rs:1439-1503"""

    def wait(self):
        return TypeConverter.icicle_to_g1_projective(self._finish())


class G2MsmHandle(_Handle):
    def wait(self):
        return TypeConverter.icicle_to_g2_projective(self._finish())


class BatchMsmHandle(_Handle):
    def batch_size(self) -> int:
        return self._result.shape[0]          # PinnedArray and ndarray both carry .shape

    def wait(self):
        res = self._finish()
        return [TypeConverter.icicle_to_g1_projective(r) for r in res]


class GpuMsmContext:
    def __init__(self, device_id: int = 0, window: int = 0):
        try:
            ensure_backend_loaded()
            set_device(device_id)
        except Exception as e:  # noqa: BLE001
            raise MsmError(f"backend init failed: {e}") from e
        self.device_id = device_id
        self.window = window          # MIDNIGHT_MSM_WINDOW as a plain parameter (core/config.rs), 0 = auto

    # -- uploads -----------------------------------------------------------
    def upload_g1_bases(self, points) -> PrecomputedBases:
        pts = TypeConverter.g1_slice_as_icicle(points)
        return PrecomputedBases(DeviceVec.from_host(pts, G1_AFFINE_BYTES), pts.shape[0])

    def upload_g2_bases(self, points) -> DeviceVec:
        pts = TypeConverter.g2_slice_as_icicle(points)
        return DeviceVec.from_host(pts, G2_AFFINE_BYTES)

    def precompute_bases(self, bases: PrecomputedBases, factor: int) -> PrecomputedBases:
        """core/msm.rs:401-492; here a real expansion 2^(k*c*ceil(W/f)) * P_i on the device."""
        if factor <= 1:
            return bases
        n = bases.original_size()
        out = DeviceVec(n * factor, G1_AFFINE_BYTES)
        cfg = self._cfg(points_on_device=True, results_on_device=True)
        cfg.precompute_factor = factor
        self._check(L.lib().b381_g1_msm_precompute_bases(L.ptr(bases.buffer), n, C.byref(cfg), L.ptr(out)), "precompute")
        return PrecomputedBases(out, n, factor, self.window)

    def upload_g1_bases_with_precompute(self, points, factor: int) -> PrecomputedBases:
        return self.precompute_bases(self.upload_g1_bases(points), factor)

    # -- G1 ------------------------------------------------------------------
    def msm(self, scalars, points):
        """host scalars (Montgomery) x host points (Montgomery affine) (core/msm.rs:519-592)."""
        sc, pts = TypeConverter.scalar_slice_as_icicle(scalars), TypeConverter.g1_slice_as_icicle(points)
        self._same_len(sc, pts)
        res = np.zeros(18, dtype=np.uint64)
        cfg = self._cfg()
        self._check(L.lib().b381_g1_msm(L.ptr(sc), L.ptr(pts), sc.shape[0], C.byref(cfg), L.ptr(res)), "msm")
        return TypeConverter.icicle_to_g1_projective(res)

    def msm_with_device_bases(self, scalars, bases: PrecomputedBases):
        """core/msm.rs:594-682 (the KZG-commit hot path)."""
        sc = TypeConverter.scalar_slice_as_icicle(scalars)
        bases.required_size_for_scalars(sc.shape[0])
        res = np.zeros(18, dtype=np.uint64)
        cfg = self._cfg_for(bases, points_on_device=True)
        n = sc.shape[0]
        self._check(L.lib().b381_g1_msm(L.ptr(sc), L.ptr(bases.buffer), n, C.byref(cfg), L.ptr(res)), "msm")
        return TypeConverter.icicle_to_g1_projective(res)

    def msm_with_device_bases_async(self, scalars, bases: PrecomputedBases) -> MsmHandle:
        sc = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(scalars))
        bases.required_size_for_scalars(sc.shape[0])
        # host scalars go straight into the plugin call: it stages them on THIS call's stream from the stream-ordered
        # pool (no cudaMalloc/cudaFree, no device-wide sync), so two handles in flight overlap copy and compute
        st = ManagedStream.create()
        res = PinnedArray(18)          # pinned: the result copy must not block this thread (see PinnedArray)
        cfg = self._cfg_for(bases, points_on_device=True, stream=st, is_async=True)
        self._check(L.lib().b381_g1_msm(L.ptr(sc), L.ptr(bases.buffer), sc.shape[0], C.byref(cfg), L.ptr(res.array)), "msm")
        return MsmHandle(st, res, (sc, bases))

    def msm_async(self, scalars, points) -> MsmHandle:
        return self.msm_with_device_bases_async(scalars, self.upload_g1_bases(points))

    def msm_batch_with_device_bases(self, scalars_batch, bases: PrecomputedBases):
        """core/msm.rs:1179-1295: B scalar vectors sharing one base set, one backend call."""
        sc = np.ascontiguousarray(np.stack([TypeConverter.scalar_slice_as_icicle(s) for s in scalars_batch]))
        b, n = sc.shape[0], sc.shape[1]
        bases.required_size_for_scalars(n)
        res = np.zeros((b, 18), dtype=np.uint64)
        cfg = self._cfg_for(bases, points_on_device=True)
        cfg.batch_size, cfg.are_points_shared_in_batch = b, True
        self._check(L.lib().b381_g1_msm(L.ptr(sc), L.ptr(bases.buffer), n, C.byref(cfg), L.ptr(res)), "msm batch")
        return [TypeConverter.icicle_to_g1_projective(r) for r in res]

    def msm_batch_with_device_bases_async(self, scalars_batch, bases: PrecomputedBases) -> BatchMsmHandle:
        sc = np.ascontiguousarray(np.stack([TypeConverter.scalar_slice_as_icicle(s) for s in scalars_batch]))
        b, n = sc.shape[0], sc.shape[1]
        bases.required_size_for_scalars(n)
        st = ManagedStream.create()
        res = PinnedArray((b, 18))
        cfg = self._cfg_for(bases, points_on_device=True, stream=st, is_async=True)
        cfg.batch_size, cfg.are_points_shared_in_batch = b, True
        self._check(L.lib().b381_g1_msm(L.ptr(sc), L.ptr(bases.buffer), n, C.byref(cfg), L.ptr(res.array)), "msm batch")
        return BatchMsmHandle(st, res, (sc, bases))

    # -- G2 ------------------------------------------------------------------
    def g2_msm(self, scalars, points):
        sc, pts = TypeConverter.scalar_slice_as_icicle(scalars), TypeConverter.g2_slice_as_icicle(points)
        self._same_len(sc, pts)
        res = np.zeros(36, dtype=np.uint64)
        cfg = self._cfg()
        self._check(L.lib().b381_g2_msm(L.ptr(sc), L.ptr(pts), sc.shape[0], C.byref(cfg), L.ptr(res)), "g2 msm")
        return TypeConverter.icicle_to_g2_projective(res)

    def g2_msm_with_device_bases(self, scalars, bases: DeviceVec):
        sc = TypeConverter.scalar_slice_as_icicle(scalars)
        if sc.shape[0] > len(bases):
            raise MsmError("more scalars than bases")
        res = np.zeros(36, dtype=np.uint64)
        cfg = self._cfg(points_on_device=True)
        self._check(L.lib().b381_g2_msm(L.ptr(sc), L.ptr(bases), sc.shape[0], C.byref(cfg), L.ptr(res)), "g2 msm")
        return TypeConverter.icicle_to_g2_projective(res)

    def g2_msm_async(self, scalars, points) -> G2MsmHandle:
        sc = np.ascontiguousarray(TypeConverter.scalar_slice_as_icicle(scalars))
        d_pts = self.upload_g2_bases(points)
        st = ManagedStream.create()
        res = PinnedArray(36)
        cfg = self._cfg(points_on_device=True, stream=st, is_async=True)
        self._check(L.lib().b381_g2_msm(L.ptr(sc), L.ptr(d_pts), sc.shape[0], C.byref(cfg), L.ptr(res.array)), "g2 msm")
        return G2MsmHandle(st, res, (sc, d_pts))

    def warmup(self) -> float:
        """core/msm.rs:931-983: one tiny MSM to pay context/module load once; returns seconds."""
        t0 = time.perf_counter()
        one = np.array([[0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5, 0x1824B159ACC5056F]], dtype=np.uint64)
        self.msm(one, np.zeros((1, 12), dtype=np.uint64))
        return time.perf_counter() - t0

    # -- helpers -------------------------------------------------------------
    def _cfg(self, points_on_device=False, scalars_on_device=False, results_on_device=False, stream=None, is_async=False):
        cfg = L.lib().b381_default_msm_config()
        cfg.c = self.window
        cfg.are_scalars_montgomery_form = True     # midnight-curves scalars are Montgomery (core/msm.rs:639-651)
        cfg.are_points_montgomery_form = True
        cfg.are_points_on_device = points_on_device
        cfg.are_scalars_on_device = scalars_on_device
        cfg.are_results_on_device = results_on_device
        cfg.is_async = is_async
        if stream is not None:
            cfg.stream = stream.handle
        return cfg

    def _cfg_for(self, bases: PrecomputedBases, **kw):
        """config of an MSM over `bases`: precomputed tables fix the factor AND the window they were built with."""
        cfg = self._cfg(**kw)
        cfg.precompute_factor = bases.factor()
        if bases.is_precomputed():
            cfg.c = bases.window
        return cfg

    @staticmethod
    def _same_len(sc, pts):
        if sc.shape[0] != pts.shape[0]:
            raise MsmError(f"Scalar count {sc.shape[0]} != base count {pts.shape[0]}")

    @staticmethod
    def _check(code, where):
        if code != 0:
            raise MsmError(f"{where}: {L.ERROR_NAMES.get(code, code)}")


assert G1_PROJECTIVE_BYTES == 18 * 8 and G2_PROJECTIVE_BYTES == 36 * 8
