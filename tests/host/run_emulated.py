#!/usr/bin/env python
"""TEST INFRASTRUCTURE: run the parity tests of the widened rows (tests/test_gpu_widen_*.py - the same test functions the
B200 box runs) against tests/host/_build/libinfercnv_b200_emu.so, i.e. against the kernels' own source text executed by
the host emulation of tests/host/emu/cuda_runtime.h.  A separate process, because it points the ctypes loader at the
emulated library before anything is loaded; the package itself has no such switch.

    python tests/host/run_emulated.py [pytest args]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import build_emu  # noqa: E402
import pytest  # noqa: E402

from infercnv_b200 import _lib  # noqa: E402

_lib.LIB_PATH = build_emu.build()
assert _lib._lib is None

# not emulated: torch device tensors (Engine), and the full-size property test (10^7 cell-genes, sized for the GPU)
DESELECT = ["test_device_resident_states_from_the_viterbi_kernel_to_regions", "test_full_size_round_trip_and_run_count"]
args = [os.path.join(ROOT, "tests", f) for f in ("test_gpu_widen_regions.py", "test_gpu_widen_ingest.py", "test_gpu_widen_denoise.py",
                                                 "test_gpu_widen_elementwise.py")]
args += ["-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "-k", " and ".join("not " + d for d in DESELECT)] + sys.argv[1:]
sys.exit(pytest.main(args))
