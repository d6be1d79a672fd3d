#!/usr/bin/env python
"""TEST INFRASTRUCTURE: bench.py's own code path (workload construction, timed loop, e2e through the host ABI, the
max-over-ranks reduction, the JSON line) executed on the CPU at a toy size - CPU tensors and the emulated library stand
in for the GPU (see engine_emulated.py), wall-clock stamps for CUDA events, gloo for NCCL.  The NUMBERS it prints mean
nothing; what is checked is that the script runs to its JSON line with every contract key, for 1 and for N ranks
(RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* in the environment, as torchrun sets them).

    python tests/host/bench_emulated.py [bench.py arguments]
"""
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import engine_emulated  # noqa: E402

engine_emulated.make_engine()            # patches torch.cuda.* / device._stream_ptr and points the loader at the emulation
from infercnv_b200 import device  # noqa: E402

_init = device.Engine.__init__


def _engine_init(self, dev=0):
    _init(self, 0)
    self.tdev = torch.device("cpu")


device.Engine.__init__ = _engine_init
torch.cuda.is_available = lambda: True


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, *a):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = _Event
_empty = torch.empty
torch.empty = lambda *a, **k: _empty(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"})
_pg = dist.init_process_group
dist.init_process_group = lambda backend=None, **k: _pg("gloo", **{kk: v for kk, v in k.items() if kk != "device_id"})

sys.argv = [os.path.join(ROOT, "bench.py")] + (sys.argv[1:] or ["--cells", "96", "--genes", "1100", "--steps", "2", "--warmup", "3",
                                                                "--ref-sample-cells", "24"])
runpy.run_path(sys.argv[0], run_name="__main__")
