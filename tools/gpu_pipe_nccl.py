#!/usr/bin/env python3
"""Two sharded MSMs in flight over NCCL (dev tool; run under torchrun): where does the per-step collective hurt?
usage: torchrun --nproc-per-node N tools/gpu_pipe_nccl.py <log_n_total> [steps]
V0 serial steps (partial, all_gather, combine to host) | V1 two in flight, all_gather + device combine on the step's stream,
handle-style wait | V2 two in flight, partials only, ONE all_gather of all K partials and K combines at the end."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench as B
import midnight_bls12_381_cuda_b200 as M
from midnight_bls12_381_cuda_b200 import _lib as L
from midnight_bls12_381_cuda_b200 import dist as D

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
M.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = L.lib()
n = 1 << int(sys.argv[1])
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_loc = n // world
g = np.array(B.G1_GEN_MONT, dtype=np.uint64)
bases = torch.empty((n_loc, 12), dtype=torch.int64, device="cuda")
L.check(lib.b381_g1_point_series(L.ptr(g), L.ptr(g), C.c_uint64(n_loc), L.ptr(bases), None), "series")
sc = B.canonical_fr(torch, n_loc, 0xB12381 + rank)
msm = D.ShardedMsm("g1")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
parts = [torch.empty(192, dtype=torch.uint8, device="cuda") for _ in range(2)]
gath = [torch.empty((world, 192), dtype=torch.uint8, device="cuda") for _ in range(2)]
res = [torch.zeros(18, dtype=torch.int64, device="cuda") for _ in range(2)]
allp = torch.empty((K, 192), dtype=torch.uint8, device="cuda")
allg = torch.empty((world, K, 192), dtype=torch.uint8, device="cuda")
allr = torch.zeros((K, 18), dtype=torch.int64, device="cuda")


def v0():
    for _ in range(K):
        p = msm.partial(sc, bases, n_loc, out=parts[0])
        D.gather_partials_into(p, gath[0])
        msm.combine(gath[0])


def v1():
    done = [None, None]
    for i in range(K):
        j = i & 1
        if done[j] is not None:
            done[j].synchronize()
        with torch.cuda.stream(streams[j]):
            msm.partial(sc, bases, n_loc, stream=streams[j].cuda_stream, out=parts[j])
            D.gather_partials_into(parts[j], gath[j])
            msm.combine_async(gath[j], res[j], stream=streams[j].cuda_stream)
            done[j] = torch.cuda.Event()
            done[j].record(streams[j])
    torch.cuda.synchronize()


def v2():
    done = [None, None]
    for i in range(K):
        j = i & 1
        if done[j] is not None:
            done[j].synchronize()
        with torch.cuda.stream(streams[j]):
            msm.partial(sc, bases, n_loc, stream=streams[j].cuda_stream, out=allp[i])
            done[j] = torch.cuda.Event()
            done[j].record(streams[j])
    torch.cuda.synchronize()
    dist.all_gather_into_tensor(allg.view(-1), allp.view(-1))
    for i in range(K):
        msm.combine_async(allg[:, i].contiguous(), allr[i])
    torch.cuda.synchronize()


for name, fn in (("V0 serial", v0), ("V1 two in flight, per-step all_gather", v1), ("V2 two in flight, one all_gather at the end", v2)):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} GPUs, 2^{sys.argv[1]} total, K={K}: {name}: {ms.item() / K:.2f} ms per step", flush=True)
if rank == 0:
    print("V1 == V2 result:", bool((allr[K - 1].cpu() == res[(K - 1) & 1].cpu()).all()))
dist.destroy_process_group()
