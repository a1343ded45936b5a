"""include/b381.hpp (C++ twin of the reference's Rust core/ API): compiles, links against the C ABI,
fails loudly without a device (CPU suite) and reproduces the reference's Rust unit tests on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory, b381):
    out = str(tmp_path_factory.mktemp("cpp") / "cpp_api_test")
    lib = os.path.join(ROOT, "midnight_bls12_381_cuda_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", "-o", out, f"{ROOT}/tests/host/cpp_api_test.cpp",
                    f"-L{lib}", "-lb381_cuda", f"-Wl,-rpath,{lib}"], check=True)
    return out


def test_cpp_header_links_and_has_no_fallback(exe):
    r = subprocess.run([exe, "link"], capture_output=True, text=True)
    assert r.returncode == 0 and "cpp api link ok" in r.stdout, r.stderr


@pytest.mark.gpu
def test_cpp_api_on_gpu(exe, cuda):
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "cpp api ok" in r.stdout, r.stdout + r.stderr
