"""bench.py's reference arm (the one leg that runs without a GPU): ONE JSON line on stdout with the keys the driver
reads, the CPU port timed on a bounded sample, rank != 0 silent under torchrun.  The GPU arm's line is produced on the
B200 (profiles/r02h_bench_*.json); here only its key set is compared with the committed line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--cpu-log-n", "10", *args], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_line():
    r = run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                      # ONE JSON line, nothing else on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "g1_msm_2^24_points_per_s" and d["unit"] == "points/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "2^10" in d["cpu_baseline"]["sample"]
    assert d["config"]["workload"].startswith("G1 MSM n=2^24")


def test_reference_arm_other_ranks_are_silent():
    r = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == "", (r.returncode, r.stdout[-500:], r.stderr[-500:])


def test_gpu_arm_line_has_the_contract_keys():
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02h_bench_plain.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "gpu_launches"} <= set(line)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] == (1 << 24) * 32 and line["gpu_launches"] > 0
    assert line["result_check"] == "ok" and line["ntt"]["result_check"] == "ok"
