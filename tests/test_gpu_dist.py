"""Multi-GPU building blocks checked on ONE GPU by running every rank's share in turn
(B200_PROFILING.md: emulate ranks on one device, never as concurrent waiting kernels)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyref as P
from vectors import fr_ints

pytestmark = pytest.mark.gpu


def test_sharded_msm_equals_single(cuda, b381, oracle):
    from midnight_bls12_381_cuda_b200 import dist as D
    n = 1 << 14
    bases = oracle.gen_series(1, [5, 0, 0, 0], [9, 0, 0, 0], n)
    sc = oracle.random_fr(77, n)
    exp = oracle.msm(1, sc, bases).tobytes()
    d_b = cuda.from_numpy(bases.view(np.int64)).cuda()
    d_s = cuda.from_numpy(sc.view(np.int64)).cuda()
    for world in (1, 2, 8):
        m = D.ShardedMsm("g1")
        parts = []
        for rank in range(world):
            b, e = D.shard_range(n, rank, world)
            parts.append(m.partial(d_s[b:e], d_b[b:e], e - b, scalars_mont=False))
        assert m.combine(parts).tobytes() == exp, world
    # G2, 2 shards
    n2 = 1 << 9
    bases2 = oracle.gen_series(2, [5, 0, 0, 0], [9, 0, 0, 0], n2)
    exp2 = oracle.msm(2, sc[:n2], bases2).tobytes()
    d_b2 = cuda.from_numpy(bases2.view(np.int64)).cuda()
    m2 = D.ShardedMsm("g2")
    parts = [m2.partial(d_s[b:e], d_b2[b:e], e - b, scalars_mont=False) for b, e in (D.shard_range(n2, r, 2) for r in range(2))]
    assert m2.combine(parts).tobytes() == exp2


def test_sharded_msm_two_in_flight(cuda, b381, oracle):
    """Two sharded MSMs in flight the way bench.py / tools/gpu_pipe_nccl.py issue them: async partials on two streams
    written into ONE [k, world, 192] tensor, then stream-ordered combines with a DEVICE result
    (b381_g1_msm_combine(result_on_device = true) does not synchronise) -- every result equals the oracle's."""
    from midnight_bls12_381_cuda_b200 import dist as D
    n, world, k = 1 << 13, 4, 5
    bases = oracle.gen_series(1, [5, 0, 0, 0], [9, 0, 0, 0], n)
    d_b = cuda.from_numpy(bases.view(np.int64)).cuda()
    scs = [oracle.random_fr(200 + i, n) for i in range(k)]
    d_ss = [cuda.from_numpy(s.view(np.int64)).cuda() for s in scs]
    m = D.ShardedMsm("g1")
    streams = [cuda.cuda.Stream(), cuda.cuda.Stream()]
    parts = cuda.empty((k, world, 192), dtype=cuda.uint8, device="cuda")
    out = cuda.zeros((k, 18), dtype=cuda.int64, device="cuda")
    cuda.cuda.synchronize()
    for i in range(k):
        s = streams[i & 1]
        with cuda.cuda.stream(s):
            for rank in range(world):
                b, e = D.shard_range(n, rank, world)
                m.partial(d_ss[i][b:e], d_b[b:e], e - b, scalars_mont=False, stream=s.cuda_stream, out=parts[i, rank])
            m.combine_async(parts[i], out[i], stream=s.cuda_stream)
    cuda.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint64)
    for i in range(k):
        assert got[i].tobytes() == oracle.msm(1, scs[i], bases).tobytes(), i
    # the gather helper into one preallocated tensor (world = 1 without a process group: a plain copy)
    into = cuda.zeros((1, 192), dtype=cuda.uint8, device="cuda")
    assert D.gather_partials_into(parts[0, 0], into) is into and bool((into[0] == parts[0, 0]).all())


@pytest.mark.parametrize("log_n,world", [(12, 2), (14, 4), (16, 8), (18, 8)])
def test_fourstep_ntt_equals_single_gpu(cuda, b381, oracle, log_n, world):
    import midnight_bls12_381_cuda_b200 as M
    from midnight_bls12_381_cuda_b200 import dist as D
    lib = b381.lib()
    ctx = M.GpuNttContext(24)
    n = 1 << log_n
    x = oracle.random_fr(log_n * 10 + world, n)
    sh = D.fourstep_shape(log_n, world)
    for direction in (0, 1):
        # single-GPU reference: kNR
        ref = cuda.from_numpy(x.view(np.int64)).cuda()
        ctx.ntt_on_device(ref.data_ptr(), direction, size=n, ordering=M.ntt.kNR)
        # every rank's column block, step 1
        locs = []
        for rank in range(world):
            loc = cuda.from_numpy(np.ascontiguousarray(D.column_block_of(x, log_n, rank, world)).view(np.int64)).cuda()
            assert lib.b381_ntt_dist_columns(b381.ptr(loc), log_n, sh["log_g"], rank, sh["a"], direction, None) == 0
            locs.append(loc)
        outs = []
        for rank in range(world):
            def a2a(recv, send, rank=rank):
                for r in range(world):
                    recv[r] = locs[r].reshape(world, -1, 4)[rank]
            rows = D.exchange_rows(locs[rank], log_n, world, a2a)
            dn = D.DistributedNtt(log_n)
            dn._rows(rows, 1 << sh["lo"], sh["rows_per_rank"], direction, 1)
            if direction == 1:
                dn._scale_pow2_inv(rows, sh["a"])
            outs.append(rows)
        got = cuda.cat(outs)
        assert cuda.equal(got, ref), (log_n, world, direction)


@pytest.mark.parametrize("log_n,world", [(12, 2), (16, 4), (18, 8)])
def test_fourstep_ntt_fused_exchange_equals_single_gpu(cuda, b381, oracle, log_n, world):
    """b381_ntt_dist_columns_p2p: the last column pass stores straight into the row buffers of the owning ranks
    (here: `world` buffers in one process standing in for the IPC-mapped peer buffers), 256-bit stores, tile order
    rotated by rank.  Afterwards every buffer must hold that rank's rows ready for the row transforms."""
    import midnight_bls12_381_cuda_b200 as M
    from midnight_bls12_381_cuda_b200 import dist as D
    lib = b381.lib()
    ctx = M.GpuNttContext(24)
    n = 1 << log_n
    x = oracle.random_fr(log_n * 7 + world, n)
    sh = D.fourstep_shape(log_n, world)
    for direction in (0, 1):
        ref = cuda.from_numpy(x.view(np.int64)).cuda()
        ctx.ntt_on_device(ref.data_ptr(), direction, size=n, ordering=M.ntt.kNR)
        rows = [cuda.zeros((sh["local"], 4), dtype=cuda.int64, device="cuda") for _ in range(world)]
        peers = (C.c_void_p * world)(*[r.data_ptr() for r in rows])
        for rank in range(world):
            loc = cuda.from_numpy(np.ascontiguousarray(D.column_block_of(x, log_n, rank, world)).view(np.int64)).cuda()
            assert lib.b381_ntt_dist_columns_p2p(b381.ptr(loc), log_n, sh["log_g"], rank, sh["a"], direction, peers, None) == 0
        dn = D.DistributedNtt(log_n)
        for rank in range(world):
            dn._rows(rows[rank], 1 << sh["lo"], sh["rows_per_rank"], direction, 1)
            if direction == 1:
                dn._scale_pow2_inv(rows[rank], sh["a"])
        assert cuda.equal(cuda.cat(rows), ref), (log_n, world, direction)
    # more than 8 GPUs do not fit ntt_pass_params::peer_out
    assert lib.b381_ntt_dist_columns_p2p(b381.ptr(rows[0]), 20, 4, 0, 10, 0, peers, None) != 0
