"""Real NCCL, world_size 2: the two exchanges of the path (MSM partial all_gather, four-step NTT all_to_all) and the
round-robin commit batch, run through bench.py's multi-GPU legs at a small size and checked byte for byte against the
single-GPU results (bench.py multi_gpu_legs).  Skipped on a one-GPU box; the host index logic is covered on CPU with
gloo in tests/test_dist.py and the per-rank kernels on one GPU in tests/test_gpu_dist.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_parity(cuda):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--log-n", "20", "--dist-ntt-log", "20", "--no-cpu-baseline"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2
    assert line["result_check"].startswith("ok"), line["result_check"]
    assert line["ntt_fourstep"]["result_check"].startswith("ok"), line["ntt_fourstep"]
    assert line["plonk_commit_round"]["result_check"].startswith("ok"), line["plonk_commit_round"]
