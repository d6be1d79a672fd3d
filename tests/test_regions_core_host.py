"""The per-element decisions of the CNV region kernels (infercnv_b200/csrc/icnv_regions_core.h: state slots, packed
byte counters with the flush-before-256 rule, modal state with ties to the smallest state, the region-opening
predicate) are written for host and device; here the same text is compiled with g++ and checked against plain
counting and a literal walk of .define_cnv_gene_regions (R/inferCNV_HMM.R:977-1058)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None and not os.path.exists("/usr/bin/g++"), reason="no host compiler")
def test_regions_core_host_build(tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    exe = str(tmp_path / "regions_core_check")
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "infercnv_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "regions_core_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout
