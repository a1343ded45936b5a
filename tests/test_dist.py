"""Host-side logic of the multi-GPU path with world_size 2 on CPU (gloo): shard ranges, partial
gather order, and the four-step NTT row exchange (index arithmetic only -- the field arithmetic of
the distributed NTT is checked on the GPU in test_gpu_dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from midnight_bls12_381_cuda_b200 import dist as D


def test_shard_ranges_cover():
    for n in (0, 1, 7, 1 << 16, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1


def test_fourstep_shape():
    sh = D.fourstep_shape(26, 8)
    assert sh == {"log_g": 3, "a": 13, "lo": 13, "L": 1024, "rows_per_rank": 1024, "local": 1 << 23}
    sh = D.fourstep_shape(24, 2)
    assert sh["a"] + sh["lo"] == 24 and sh["L"] * 2 == 1 << sh["lo"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, log_n, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # partial gather keeps rank order
        part = torch.full((192,), rank, dtype=torch.uint8)
        got = D.gather_partials(part)
        assert [int(g[0]) for g in got] == list(range(world))
        into = torch.empty((world, 192), dtype=torch.uint8)
        assert D.gather_partials_into(part, into) is into and into[:, 0].tolist() == list(range(world)) \
            and into[:, 191].tolist() == list(range(world))
        # row exchange: tag every element with its global flat index
        sh = D.fourstep_shape(log_n, world)
        x = np.arange(1 << log_n, dtype=np.int64)
        local = torch.from_numpy(np.ascontiguousarray(D.column_block_of(x, log_n, rank, world)))

        def a2a(recv, send):
            try:
                dist.all_to_all_single(recv, send)
            except RuntimeError:          # older gloo builds: emulate with all_gather
                bufs = [torch.empty_like(send) for _ in range(world)]
                dist.all_gather(bufs, send.contiguous())
                for r in range(world):
                    recv[r] = bufs[r][rank]
        rows = D.exchange_rows(local, log_n, world, a2a)
        R = sh["rows_per_rank"]
        exp = x.reshape(1 << sh["a"], 1 << sh["lo"])[rank * R:(rank + 1) * R].reshape(-1)
        assert np.array_equal(rows.numpy(), exp)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("log_n", [8, 11])
def test_world2_gloo(log_n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), log_n, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
