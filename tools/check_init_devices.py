#!/usr/bin/env python
"""One host process on several GPUs through the C ABI (icnv_init_devices, what the R shim's infercnvb200_init_devices() calls):
the streaming entry points must return the same bytes as on one GPU, from pageable host memory.  Prints the end-to-end time of
the fused call at the c2 size for 1 and for all GPUs.     python tools/check_init_devices.py   (on a box with >= 2 GPUs)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import api  # noqa: E402
from infercnv_b200.hmm import CNV_LEVELS, get_HMM  # noqa: E402
from oracle import oracle as orc  # noqa: E402

G, C = 10000, int(os.environ.get("ICNV_CHECK_CELLS", "10000"))
cs, cl = bench.chr_layout(G)
refs = [np.asarray(g, dtype=np.int32) for g in bench.ref_groups_global(C)]
X = orc.synth(G, cs, cl, np.arange(C), C, bench.SEED, nthreads=bench.usable_cpus())       # pageable NumPy memory
Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, bench.I6_MEAN, bench.I6_SD)}, 1e-6)


def run():
    Y = np.empty_like(X, order="F")
    S = np.empty((G, C), dtype=np.uint8, order="F")
    api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)   # warm-up (allocations)
    t0 = time.perf_counter()
    api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)
    dt = time.perf_counter() - t0
    Yb = api.smooth_block(X, cs, cl, refs)
    Sv = np.empty((G, C), dtype=np.uint8, order="F")
    api.viterbi(Yb, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Sv)
    return Y, S, Yb, Sv, dt


api.init(0)
Y1, S1, Yb1, Sv1, t1 = run()
n = api.init_devices(None)
Yn, Sn, Ybn, Svn, tn = run()
ok = np.array_equal(Y1, Yn) and np.array_equal(S1, Sn) and np.array_equal(Yb1, Ybn) and np.array_equal(Sv1, Svn) and np.array_equal(Y1, Yb1)
print(f"[check_init_devices] {n} GPUs from one process, {C} cells x {G} genes, pageable host memory: fused call {t1 * 1e3:.1f} ms on 1 GPU, "
      f"{tn * 1e3:.1f} ms on {n}; outputs {'BITWISE EQUAL' if ok else 'MISMATCH'}")
sys.exit(0 if ok and n >= 1 else 1)
