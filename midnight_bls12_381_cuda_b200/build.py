"""In-tree build of the CUDA libraries for sm_100a (nvcc cross-compiles without a GPU).

Artefacts (all under midnight_bls12_381_cuda_b200/lib/, git-ignored, shipped by gpurun):
  libb381_cuda.so                                  kernels + flat C ABI (include/b381.h)
  libicicle_backend_cuda_field_bls12_381.so        ICICLE registration: NTT + vecops
  libicicle_backend_cuda_curve_bls12_381.so        ICICLE registration: G1/G2 MSM (+ g2 registry)
  libicicle_backend_cuda_device.so                 ICICLE "CUDA" DeviceAPI
(names and split follow bls12-381/CMakeLists.txt:158-189 / scripts/icicle_install.sh:21-22)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "build")
ROOT = os.path.dirname(HERE)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", CSRC, "-I", os.path.join(ROOT, "include"),
]

CORE_SRCS = ["msm_g1.cu", "msm_g2.cu", "msm_pair.cu", "msm_sort.cu", "msm_tail.cu", "ntt.cu", "vecops.cu", "devapi.cu", "probes.cu", "pointgen.cu", "points.cu", "point_mul.cu"]
ICICLE_FIELD_SRCS = ["icicle/field_api.cu"]
ICICLE_CURVE_SRCS = ["icicle/curve_api.cu", "icicle/g2_registry.cu"]
ICICLE_DEVICE_SRCS = ["icicle/device_api.cu"]


def _deps_mtime() -> float:
    m = 0.0
    for d, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cuh", ".h", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    m = max(m, os.path.getmtime(os.path.join(ROOT, "include", "b381.h")))
    return m


def _deps_of(obj: str) -> list[str] | None:
    """headers recorded by nvcc -MD for this object (None = no record yet)."""
    dep = obj + ".d"
    if not os.path.exists(dep):
        return None
    txt = open(dep).read().replace("\\\n", " ")
    parts = txt.split(":", 1)
    if len(parts) < 2:
        return None
    return [t for t in parts[1].split() if t.startswith(ROOT)]


def _compile(src: str, hdr_mtime: float, verbose: bool, defines: tuple = (), tag: str = "") -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src.replace("/", "_") + tag + ".o")
    if os.path.exists(obj):
        deps = _deps_of(obj)
        newest = hdr_mtime if deps is None else max(
            [os.path.getmtime(d) if os.path.exists(d) else float("inf") for d in deps] + [0.0])
        if os.path.getmtime(obj) > max(os.path.getmtime(path), newest):
            return obj
    cmd = [NVCC] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-MD", "-MF", obj + ".d", "-c", path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def _link(name: str, objs: list[str], extra: list[str] | None = None) -> str:
    out = os.path.join(LIB, name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if os.path.exists(out) and all(os.path.getmtime(out) > os.path.getmtime(o) for o in objs):
        return out
    cmd = [NVCC, "-shared", "-o", out] + objs + ["-Xcompiler", "-fPIC"] + (extra or [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed for {name}:\n{r.stdout}\n{r.stderr}")
    return out


def generate_field_header() -> None:
    gen = os.path.join(CSRC, "gen", "gen_field.py")
    out = os.path.join(CSRC, "field_ptx.cuh")
    if (not os.path.exists(out)) or os.path.getmtime(out) < max(
            os.path.getmtime(gen), os.path.getmtime(os.path.join(CSRC, "gen", "ptxir.py"))):
        subprocess.run([sys.executable, gen], check=True, capture_output=True)


def build(verbose: bool = False, icicle: bool = True) -> dict[str, str]:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    generate_field_header()
    hdr = _deps_mtime()
    groups = {"core": CORE_SRCS}
    if icicle:
        groups.update(field=ICICLE_FIELD_SRCS, curve=ICICLE_CURVE_SRCS, device=ICICLE_DEVICE_SRCS)
    all_srcs = [s for g in groups.values() for s in g if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=min(8, len(all_srcs))) as ex:
        objs = dict(zip(all_srcs, ex.map(lambda s: _compile(s, hdr, verbose), all_srcs)))
    core = [objs[s] for s in CORE_SRCS if s in objs]
    out = {"core": _link("libb381_cuda.so", core)}
    if icicle:
        def pick(names):
            return [objs[s] for s in names if s in objs]
        # field lib: NTT + vecops kernels + registration; curve lib: MSM kernels + registration
        if pick(ICICLE_FIELD_SRCS):
            out["field"] = _link("libicicle_backend_cuda_field_bls12_381.so",
                                 pick(ICICLE_FIELD_SRCS) + [objs[s] for s in ("ntt.cu", "vecops.cu") if s in objs])
            # same library for a real ICICLE v4 install, whose VecOpsConfig carries batch_size / columns_batch
            # (csrc/icicle/icicle_abi.h): lib/upstream_v4/, same file name
            v4 = _compile(ICICLE_FIELD_SRCS[0], hdr, verbose, ("B381_ICICLE_UPSTREAM_VECOPS",), "_v4")
            out["field_v4"] = _link(os.path.join("upstream_v4", "libicicle_backend_cuda_field_bls12_381.so"),
                                    [v4] + [objs[s] for s in ("ntt.cu", "vecops.cu") if s in objs])
        if pick(ICICLE_CURVE_SRCS):
            out["curve"] = _link("libicicle_backend_cuda_curve_bls12_381.so",
                                 pick(ICICLE_CURVE_SRCS) + [objs["points.cu"], objs["msm_g1.cu"], objs["msm_g2.cu"], objs["msm_pair.cu"], objs["msm_sort.cu"], objs["msm_tail.cu"]])
        if pick(ICICLE_DEVICE_SRCS):
            out["device"] = _link("libicicle_backend_cuda_device.so", pick(ICICLE_DEVICE_SRCS))
    return out


if __name__ == "__main__":
    for k, v in build(verbose="-v" in sys.argv).items():
        print(k, v)
