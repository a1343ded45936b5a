"""Multi-GPU execution of the hot path on one NVSwitch box: one process per GPU, torch.distributed
(NCCL) for the only two exchanges the path has.  The reference is single-GPU everywhere
(core/config.rs:528-531, core/msm.rs:284); this is the north_star's scale-out.

  * MSM  -- contiguous point ranges, one XYZZ partial (192 B G1 / 384 B G2) per GPU, all_gather of the
            partials, one combine + inversion on every rank (cheap) or rank 0 only.
  * NTT  -- four-step over column blocks: local upper stages whose LAST pass stores straight into the owning
            GPU's row buffer over NVLink peer memory (b381_ntt_dist_columns_p2p: exchange and transpose fused into
            the kernel's epilogue), two 4-byte all_reduce barriers, local row NTTs.  The NCCL path (one
            all_to_all of row blocks + transpose) stays as `fused=False`.  Output = global kNR order,
            block-distributed.

torch is plumbing here (device buffers + collectives); every field/curve operation happens in the
CUDA library.  The index arithmetic below is pure host logic and is unit-tested with gloo on CPU
(tests/test_dist.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

XYZZ_BYTES = {"g1": 192, "g2": 384}


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous [begin, end) of rank's points; remainders go to the low ranks."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_partials(partial, dist_module=None, group=None):
    """all_gather of one fixed-size byte tensor per rank -> list ordered by rank."""
    import torch
    import torch.distributed as dist
    d = dist_module or dist
    world = d.get_world_size(group) if d.is_initialized() else 1
    if world == 1:
        return [partial]
    out = [torch.empty_like(partial) for _ in range(world)]
    d.all_gather(out, partial, group=group)
    return out


def gather_partials_into(partial, out, dist_module=None, group=None):
    """same exchange into ONE preallocated [world, bytes] tensor (no per-rank output tensors, no concatenation before
    the combine): NCCL all_gather_into_tensor.  Returns `out`."""
    import torch.distributed as dist
    d = dist_module or dist
    world = d.get_world_size(group) if d.is_initialized() else 1
    if world == 1:
        out.view(-1)[:partial.numel()].copy_(partial.view(-1))
        return out
    d.all_gather_into_tensor(out.view(-1), partial.view(-1), group=group)
    return out


class ShardedMsm:
    """Point-range sharded MSM.  Each rank calls `run` with ITS shard resident on its GPU."""

    def __init__(self, curve: str = "g1", window: int = 0):
        assert curve in XYZZ_BYTES
        self.curve, self.window = curve, window
        self._lib = L.lib()

    def partial(self, scalars_dev, bases_dev, n_local: int, scalars_mont: bool = True, stream=None, out=None):
        import torch
        cfg = self._lib.b381_default_msm_config()
        cfg.c = self.window
        cfg.are_scalars_on_device = cfg.are_points_on_device = True
        cfg.are_scalars_montgomery_form = scalars_mont
        cfg.are_points_montgomery_form = True
        cfg.is_async = True
        if stream is not None:
            cfg.stream = C.c_void_p(stream)
        if out is None:
            out = torch.empty(XYZZ_BYTES[self.curve], dtype=torch.uint8, device="cuda")
        fn = self._lib.b381_g1_msm_partial if self.curve == "g1" else self._lib.b381_g2_msm_partial
        L.check(fn(L.ptr(scalars_dev), L.ptr(bases_dev), n_local, C.byref(cfg), L.ptr(out)), "msm_partial")
        return out

    def combine_async(self, partials, out_dev, stream=None):
        """sum of XYZZ partials -> ICICLE standard-form projective words in `out_dev` (device int64[18] / [36]),
        stream-ordered on `stream`: no host synchronisation, so several sharded MSMs can be in flight."""
        count = partials.numel() * partials.element_size() // XYZZ_BYTES[self.curve]
        fn = self._lib.b381_g1_msm_combine if self.curve == "g1" else self._lib.b381_g2_msm_combine
        L.check(fn(L.ptr(partials), count, C.c_void_p(stream) if stream is not None else None, True, L.ptr(out_dev)),
                "msm_combine")
        return out_dev

    def combine(self, partials) -> np.ndarray:
        """sum of XYZZ partials -> ICICLE standard-form projective bytes (host)."""
        import torch
        if isinstance(partials, torch.Tensor):               # one [world, bytes] tensor (gather_partials_into)
            allp, count = partials, partials.numel() * partials.element_size() // XYZZ_BYTES[self.curve]
        else:
            allp, count = torch.cat(list(partials)).contiguous(), len(partials)
        k = 18 if self.curve == "g1" else 36
        res = np.zeros(k, dtype=np.uint64)
        fn = self._lib.b381_g1_msm_combine if self.curve == "g1" else self._lib.b381_g2_msm_combine
        L.check(fn(L.ptr(allp), count, None, False, L.ptr(res)), "msm_combine")
        return res

    def run(self, scalars_dev, bases_dev, n_local: int, scalars_mont: bool = True, group=None):
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if getattr(self, "_gathered", None) is None or self._gathered.shape[0] != world:
            self._gathered = torch.empty((world, XYZZ_BYTES[self.curve]), dtype=torch.uint8, device="cuda")
            self._part = torch.empty(XYZZ_BYTES[self.curve], dtype=torch.uint8, device="cuda")
        part = self.partial(scalars_dev, bases_dev, n_local, scalars_mont, out=self._part)
        return self.combine(gather_partials_into(part, self._gathered, group=group))


# ----------------------------------------------------------------------------- four-step NTT
def fourstep_shape(log_n: int, world: int) -> dict:
    """Split 2^log_n = 2^a (upper stages, done on column blocks) x 2^lo (row transforms)."""
    log_g = world.bit_length() - 1
    assert 1 << log_g == world, "world size must be a power of two"
    lo = (log_n + 1) // 2
    lo = max(lo, log_g + 2)
    a = log_n - lo
    assert a >= log_g, "transform too small for this many GPUs"
    return {"log_g": log_g, "a": a, "lo": lo, "L": (1 << lo) // world, "rows_per_rank": (1 << a) // world,
            "local": (1 << log_n) // world}


def column_block_of(x, log_n: int, rank: int, world: int):
    """host helper: the local array of `rank` = x viewed as [2^a][2^lo], columns [rank*L, (rank+1)*L)."""
    sh = fourstep_shape(log_n, world)
    m = x.reshape(1 << sh["a"], 1 << sh["lo"], *x.shape[1:])
    return m[:, rank * sh["L"]:(rank + 1) * sh["L"]].reshape(sh["local"], *x.shape[1:])


def exchange_rows(local, log_n: int, world: int, all_to_all):
    """The one exchange of the four-step NTT.  `local` is [2^a][L] (this rank's columns, all rows);
    afterwards the rank owns rows [rank*R, (rank+1)*R) complete: returns [R][2^lo].
    `all_to_all(recv, send)` is torch.distributed.all_to_all_single (or an emulation in tests)."""
    sh = fourstep_shape(log_n, world)
    R, Lc = sh["rows_per_rank"], sh["L"]
    tail = tuple(local.shape[1:])
    send = local.reshape(world, R * Lc, *tail)          # rows of destination s are contiguous
    recv = send.new_empty(send.shape)
    all_to_all(recv, send)
    # recv[r] = [R][L] block holding columns of source rank r  ->  [R][world][L] -> [R][2^lo]
    perm = (1, 0, 2) + tuple(range(3, 3 + len(tail)))
    return recv.reshape(world, R, Lc, *tail).permute(*perm).reshape(R << sh["lo"], *tail).contiguous()


class DistributedNtt:
    """Forward/inverse NTT of length 2^log_n spread over `world` GPUs.
    Input: this rank's column block (see `column_block_of`), shape [2^log_n / world, 4] int64/uint64 on
    the rank's GPU.  Output: rows [rank*R, (rank+1)*R) of the [2^a][2^lo] matrix whose flat global
    position I holds X[bitrev(I)] -- i.e. the global kNR result, block-distributed."""

    def __init__(self, log_n: int, group=None, fused: bool | None = None):
        """`fused` (default: on for 2..8 GPUs): the last column pass stores straight into the row buffers of the
        owning GPUs over NVLink peer memory (b381_ntt_dist_columns_p2p), so the all_to_all and the transpose pass
        are replaced by two 4-byte all_reduce barriers.  `fused=False` keeps the NCCL all_to_all path."""
        import torch.distributed as dist
        self.log_n, self.group = log_n, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shape = fourstep_shape(log_n, self.world) if self.world > 1 else None
        self._lib = L.lib()
        self.fused = (1 < self.world <= 8) if fused is None else (fused and 1 < self.world <= 8)
        self._peers = None

    def _setup_peers(self):
        """one peer-visible row buffer per rank (cudaMalloc + CUDA IPC), every rank maps all of them"""
        import torch
        import torch.distributed as dist
        nbytes = self.shape["local"] * 32
        mine, handle = C.c_void_p(), (C.c_ubyte * 64)()
        L.check(self._lib.b381_ipc_alloc(C.c_size_t(nbytes), C.byref(mine), handle), "ipc_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        self._peers = (C.c_void_p * self.world)()
        ok = 1
        for r, hb in enumerate(handles):
            if r == self.rank:
                self._peers[r] = mine.value
                continue
            q = C.c_void_p()
            if self._lib.b381_ipc_open((C.c_ubyte * 64).from_buffer_copy(hb), C.byref(q)) != 0:
                ok = 0                      # no peer access between these two GPUs
                break
            self._peers[r] = q.value
        # every rank must take the same path: one rank without peer access sends all of them to the NCCL exchange
        agree = torch.tensor([ok], device="cuda")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if int(agree.item()) == 0:
            self.fused, self._peers = False, None
            return
        self._flag = torch.zeros(1, device="cuda")

        class _Rows:       # zero-copy torch view of the row buffer
            __cuda_array_interface__ = {"shape": (self.shape["local"], 4), "typestr": "<i8", "data": (mine.value, False),
                                        "version": 3, "strides": None}
        self._rows_view = torch.as_tensor(_Rows(), device="cuda")

    def _barrier(self):
        """stream-ordered barrier across the ranks (no host synchronisation)"""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def _rows(self, data, size, batch, direction, ordering):
        cfg = self._lib.b381_default_ntt_config()
        cfg.batch_size, cfg.ordering = batch, ordering
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = True
        L.check(self._lib.b381_ntt(L.ptr(data), size, direction, C.byref(cfg), L.ptr(data)), "ntt rows")

    def forward(self, local, direction: int = 0, marks=None):
        """`marks`, if given, is called with a label at the boundaries columns | exchange | rows (bench.py records
        CUDA events there)."""
        import torch.distributed as dist
        mark = marks or (lambda _label: None)
        if self.world == 1:
            self._rows(local, 1 << self.log_n, 1, direction, 1)        # kNR
            return local
        sh = self.shape
        if self.fused:
            if self._peers is None:
                self._setup_peers()
        if self.fused:
            rows = self._rows_view
            mark("begin")
            self._barrier()                # every rank is done with the previous contents of its row buffer
            L.check(self._lib.b381_ntt_dist_columns_p2p(L.ptr(local), self.log_n, sh["log_g"], self.rank, sh["a"], direction,
                                                        self._peers, None), "ntt_dist_columns_p2p")
            mark("columns")
            self._barrier()                # every rank's column pass, i.e. every remote store, has completed
            mark("exchange")
            self._rows(rows, 1 << sh["lo"], sh["rows_per_rank"], direction, 1)
            if direction == 1:
                self._scale_pow2_inv(rows, sh["a"])
            mark("rows")
            return rows
        mark("begin")
        L.check(self._lib.b381_ntt_dist_columns(L.ptr(local), self.log_n, sh["log_g"], self.rank, sh["a"], direction, None),
                "ntt_dist_columns")
        mark("columns")
        rows = exchange_rows(local, self.log_n, self.world,
                             lambda recv, send: dist.all_to_all_single(recv, send, group=self.group))
        mark("exchange")
        self._rows(rows, 1 << sh["lo"], sh["rows_per_rank"], direction, 1)
        if direction == 1:
            # the row transforms scaled by 2^-lo; the remaining 2^-a is one scalar multiplication
            self._scale_pow2_inv(rows, sh["a"])
        mark("rows")
        return rows

    def _scale_pow2_inv(self, data, a: int):
        r = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
        v = pow(pow(2, a, r), -1, r) * (1 << 256) % r
        s = np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        cfg = self._lib.b381_default_vecops_config()
        cfg.is_b_on_device = cfg.is_result_on_device = True
        n = data.numel() // 4
        L.check(self._lib.b381_scalar_mul_vec(L.ptr(s), L.ptr(data), C.c_uint64(n), C.byref(cfg), L.ptr(data)), "scale")
