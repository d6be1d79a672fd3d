#!/usr/bin/env python
"""TEST INFRASTRUCTURE: build tests/host/_build/libinfercnv_b200_emu.so - the byte / index kernels of the library
(icnv_regions.cu, icnv_ingest.cu, icnv_reduce.cu) and the host entry points of icnv_api.cu compiled by g++ from the SAME
source text against the execution-model emulation in tests/host/emu/cuda_runtime.h.  The only transformation is
syntactic: `kernel<<<grid, block, smem, stream>>>(args);` becomes `EMU_LAUNCH(kernel, (grid, block, smem, stream), (args));`.
The FP64 pipeline kernels (icnv_smooth.cu, icnv_viterbi.cu, icnv_median_filter.cu, icnv_synth.cu: inline PTX, TMA,
mbarriers) are not emulated; their entry points are stubs that return ICNV_E_UNSUPPORTED.

Used by tests/test_emulated_kernels.py only.  The package never loads this library."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "infercnv_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libinfercnv_b200_emu.so")
EMULATED = ["icnv_api.cu", "icnv_regions.cu", "icnv_ingest.cu", "icnv_reduce.cu"]

LAUNCH = re.compile(r"(?P<name>[A-Za-z_]\w*(?:<[^<>();]*>)?)\s*<<<(?P<cfg>.*?)>>>\s*\((?P<args>.*?)\);", re.S)


def transform(text: str) -> tuple[str, int]:
    return LAUNCH.subn(lambda m: f"EMU_LAUNCH({m['name']}, ({m['cfg']}), ({m['args']}));", text)


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "emu", f) for f in os.listdir(os.path.join(HERE, "emu"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs + [__file__]):
        return OUT
    gen = os.path.join(OUT_DIR, "src")
    shutil.rmtree(gen, ignore_errors=True)
    os.makedirs(gen)
    files = []
    for f in EMULATED:
        text, n = transform(open(os.path.join(CSRC, f)).read())
        assert "<<<" not in text, f"{f}: a launch was not recognised"
        dst = os.path.join(gen, f.replace(".cu", ".emu.cpp"))
        open(dst, "w").write(f"// generated from infercnv_b200/csrc/{f} by tests/host/build_emu.py ({n} launches rewritten)\n" + text)
        files.append(dst)
    files.append(os.path.join(HERE, "emu", "emu_stubs.cpp"))
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd = [gxx, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-D__CUDACC__", "-DICNV_EMU", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-attributes", "-Wno-unused-function", "-I", os.path.join(HERE, "emu"), "-I", CSRC, "-o", OUT, *files]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
