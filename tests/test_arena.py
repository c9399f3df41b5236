"""Host logic of the per-context scratch arena (cloops2_b200/csrc/cl2_common.cuh), compiled with g++ and run on the CPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_arena_random_alloc_free(tmp_path):
    cuda_inc = "/usr/local/cuda/include"
    if not shutil.which("g++") or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are missing")
    exe = str(tmp_path / "arena_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", cuda_inc, "-I", os.path.join(ROOT, "cloops2_b200", "csrc"),
                           os.path.join(ROOT, "tests", "arena_check.cpp"), "-o", exe, "-L/usr/local/cuda/lib64", "-lcudart"])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
