"""The per-thread window median of the CUDA median filter (infercnv_b200/csrc/icnv_median_select.cuh) is written
for host and device; here the same text is compiled with g++ and checked against a sort on random, skewed,
bimodal, tie-dominated, constant, signed-zero, extreme-magnitude and truncated windows for radius 2..5
(apply_median_filtering window_size 3..9, R/noise_reduction.R:93-113)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
def test_window_median_host_build(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "median_select_check")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "infercnv_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "median_select_check.cpp"), "-o", exe])
    out = subprocess.run([exe, "4000"], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
