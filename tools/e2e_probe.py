#!/usr/bin/env python
"""End-to-end time of the fused host call (icnv_smooth_hmm_u8_f64, pageable NumPy memory in and out) at the c3 size for several
sizes of the library's copy-thread pool and slab sizes: what bounds the host-pointer path on this box.
    python tools/e2e_probe.py [cells]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import api  # noqa: E402
from infercnv_b200.hmm import CNV_LEVELS, get_HMM  # noqa: E402
from oracle import oracle as orc  # noqa: E402

G, C = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cs, cl = bench.chr_layout(G)
refs = [np.asarray(g, dtype=np.int32) for g in bench.ref_groups_global(C)]
X = orc.synth(G, cs, cl, np.arange(C), C, bench.SEED, nthreads=bench.usable_cpus())
Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, bench.I6_MEAN, bench.I6_SD)}, 1e-6)
Y = np.empty_like(X, order="F")
S = np.empty((G, C), dtype=np.uint8, order="F")
print(json.dumps({"usable_cpus": bench.usable_cpus(), "cells": C, "bytes_in": X.nbytes, "bytes_out": Y.nbytes + S.nbytes}), flush=True)
# plain host memcpy bandwidth of this box, one thread and numpy's copy (a floor for what the staging copies can reach)
t0 = time.perf_counter()
np.copyto(Y, X)
print(json.dumps({"numpy_copy_GBps_one_thread": X.nbytes / (time.perf_counter() - t0) / 1e9}), flush=True)
for threads in ((0, 4, 16, 32) if os.environ.get("E2E_PROBE_THREADS", "1") == "1" else ()):
    api.set_host_threads(threads)
    api.reinit()
    api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)
        ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(json.dumps({"copy_threads": threads or "default", "ms": dt * 1e3, "cell_genes_per_s": G * C / dt,
                      "host_GBps_in_plus_out": (X.nbytes * 1.1 + Y.nbytes + S.nbytes) / dt / 1e9}), flush=True)
api.set_host_threads(0)
for slab in [int(v) for v in os.environ.get("E2E_PROBE_SLABS", "512,2048,4096").split(",")]:
    os.environ["ICNV_SLAB_CELLS"] = str(slab)
    api.reinit()
    api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        api.smooth_hmm(X, cs, cl, refs, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=Y, out_states=S)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"slab_cells": slab, "ms": min(ts) * 1e3, "cell_genes_per_s": G * C / min(ts)}), flush=True)
