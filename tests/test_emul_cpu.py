"""Host emulation of simple CUDA kernels (tests/emul): the UNCHANGED kernel source of csrc/ds_skinny.cu is compiled by
g++ against a stand-in <cuda_runtime.h> (every CUDA thread a std::thread, __syncthreads a barrier, warp shuffles through
an exchange buffer) and checked against naive loops — index arithmetic, the shuffle/shared-memory reduction tree and the
partial-sum layout are verified without a GPU; ThreadSanitizer looks for unsynchronised shared-memory accesses."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "emul_skinny.cpp")


def _build_and_run(tmp_path, extra):
    exe = os.path.join(tmp_path, "emul_skinny")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-DDS_EMULATE", "-I", os.path.join(HERE, "emul"), *extra, SRC,
           "-o", exe]
    subprocess.run(cmd, check=True, cwd=ROOT, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_skinny_kernels_on_the_host_emulator(tmp_path):
    res = _build_and_run(str(tmp_path), [])
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok ") == 11 and "FAIL" not in res.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_skinny_kernels_have_no_shared_memory_race(tmp_path):
    res = _build_and_run(str(tmp_path), ["-g", "-fsanitize=thread"])
    if "FATAL: ThreadSanitizer" in res.stderr and "unexpected memory mapping" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (ASLR settings)")
    assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr, res.stdout + res.stderr[-2000:]
