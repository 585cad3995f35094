"""The C ABI consumed without Python or torch in the loop: tests/abi/abi_smoke.cpp is compiled against
include/whmr_b200.h + libwhmr_b200.so + cudart and run as a separate process."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torch_free_abi_consumer(tmp_path):
    from whmr_b200 import _lib
    _lib.build()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cxx = shutil.which("g++")
    assert cxx, "g++ not found"
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    cmd = [cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "abi", "abi_smoke.cpp"), "-L", libdir, "-lwhmr_b200", "-L", os.path.join(cuda, "lib64"),
           "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ABI SMOKE OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
