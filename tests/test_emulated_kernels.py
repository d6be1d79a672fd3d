"""The library's kernels executed on the CPU from their own CUDA source text: tests/host/build_emu.py compiles every
translation unit of infercnv_b200/csrc with g++ against a host emulation of the CUDA execution model
(tests/host/emu/cuda_runtime.h: the threads of a block are fibers; __syncthreads, warp shuffles / reductions and atomics
have their CUDA semantics; a divergent barrier or a live-lock aborts; fresh device and shared memory is poisoned), with
synchronous stand-ins for the few inline-PTX copy primitives (TMA bulk copy + mbarrier, cp.async), and
tests/host/run_emulated.py runs the SAME parity tests the B200 box runs (-m gpu) against that build in a subprocess.

Test infrastructure only: it exercises the kernels' index arithmetic, barriers, shared-memory hand-offs and memory
accesses without a GPU - the fused cell pipeline, both Viterbi kernels with the certificate / re-run logic, the median
filter, and the kernels of the widened rows.  It says nothing about performance, floating-point contraction differs
from nvcc's in the last bit (the tests' tolerances are unchanged), and the package never loads the emulated library
(tests/test_capi_symbols.py)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
@pytest.mark.parametrize("order", ["forward", "reverse"])
def test_gpu_parity_tests_pass_under_the_host_emulation(order):
    """`order` = the order in which the emulation resumes the runnable threads of a block (EMU_ORDER; "random:<seed>"
    also exists): results that depend on it would mean a data race - a missing barrier or an unsynchronised hand-off
    through shared memory - in one of the kernels."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "run_emulated.py")], capture_output=True, text=True,
                       timeout=1500, env=dict(os.environ, EMU_ORDER=order))
    tail = "\n".join(r.stdout.splitlines()[-25:]) + r.stderr[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
def test_differential_fuzz_of_edge_shapes_under_the_host_emulation():
    """Random small shapes - single genes / cells, one-gene chromosomes, windows longer than a chromosome, constant and
    tie-dominated columns, groups of one - through smooth block, centring, both Viterbi arithmetics (cells and groups,
    i6 and i3), median filter and region calling, each against the oracle (tests/host/fuzz_emulated.py)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "host", "fuzz_emulated.py"), "40", "20260923"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "40 cases ok" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
@pytest.mark.parametrize("world", [1, 2])
def test_engine_and_multi_rank_path_under_the_host_emulation(world):
    """infercnv_b200/device.py (Engine) on CPU tensors against the emulated library: one rank against the oracle
    (smooth block, HMM, i3 mu/sigma, device-resident consensus and regions incl. a strided state matrix, median filter);
    two gloo ranks - the sharded smooth block with its all-gathered partial sums, the HMM, mu/sigma and the all-reduced
    region consensus - bitwise equal to the single-rank run (what tools/check_multigpu.py checks over NCCL)."""
    cmd = [sys.executable, os.path.join(ROOT, "tests", "host", "engine_emulated.py")] + (["--world", str(world)] if world > 1 else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert ("BITWISE EQUAL" in r.stdout) if world > 1 else ("equal to the oracle" in r.stdout), r.stdout[-1500:]


def test_the_package_cannot_reach_the_emulated_library():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "infercnv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".R")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "_emu" not in src and "cuda_emu" not in src and "ICNV_EMU" not in src, f
