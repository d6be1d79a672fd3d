"""The byte / index kernels of the widened rows (region calling, gene filters + sparse ingest, outlier clamp / noise
clearing, the element-wise steps) executed on the CPU from their own CUDA source text: tests/host/build_emu.py compiles
icnv_regions.cu, icnv_ingest.cu, icnv_reduce.cu and the host entry points of icnv_api.cu with g++ against a host
emulation of the CUDA execution model (tests/host/emu/cuda_runtime.h: blocks of fibers, __syncthreads, warp shuffles,
atomics; divergent barriers abort), and tests/host/run_emulated.py runs the SAME parity tests the B200 box runs
(tests/test_gpu_widen_*.py) against that build in a subprocess.

Test infrastructure only: it exercises the kernels' index arithmetic, barriers and memory accesses without a GPU; it
says nothing about performance, and the package never loads the emulated library (tests/test_capi_symbols.py)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
def test_widened_rows_pass_their_parity_tests_under_the_host_emulation():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "run_emulated.py")], capture_output=True, text=True,
                       timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:]) + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_the_package_cannot_reach_the_emulated_library():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "infercnv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "_emu" not in src and "cuda_emu" not in src and "ICNV_EMU" not in src, f
