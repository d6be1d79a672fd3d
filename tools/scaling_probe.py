#!/usr/bin/env python
"""Where the fixed cost of a small shard sits: the two hot stages timed alone at 6 250 .. 50 000 cells x 10 000 genes on one
GPU (CUDA events, 10 runs after 3 warm-ups) - the per-rank sizes of c3 at 16 .. 2 GPUs.  t(C) = f + w C: prints the least
squares f (ms) and w (ms per 1000 cells) per stage.     python tools/scaling_probe.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import dist as shard  # noqa: E402
from infercnv_b200.device import Engine  # noqa: E402
from infercnv_b200.hmm import CNV_LEVELS, get_HMM  # noqa: E402

eng = Engine(0)
G = 10000
cs, cl = bench.chr_layout(G)
Pi, delta, _, _ = get_HMM({k: {"mean": m, "sd": s} for k, m, s in zip(CNV_LEVELS, bench.I6_MEAN, bench.I6_SD)}, 1e-6)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rows = []
for C in (6250, 12500, 25000, 50000):
    refs = bench.ref_groups_global(C)
    plan = shard.plan_shards(C, refs, 1)[0]
    X = eng.synth(G, cs, cl, plan.local_cells, C, bench.SEED)
    Y = torch.empty_like(X)
    S = torch.empty((C, G), dtype=torch.uint8, device=X.device)
    rl = plan.local_ref_groups()
    t_s = timed(lambda: eng.smooth_block(X, cs, cl, rl, plan.ref_sizes, plan.max_chunks, out=Y))
    t_h = timed(lambda: eng.viterbi(Y, cs, cl, Pi, delta, bench.I6_MEAN, bench.I6_SD, out=S))
    reruns = int(eng.lib.icnv_hmm_rerun_count())
    rows.append({"cells": C, "smooth_block_ms": t_s, "hmm_ms": t_h, "hmm_reruns": reruns})
    print(json.dumps(rows[-1]), flush=True)
    del X, Y, S
c = np.array([r["cells"] for r in rows], dtype=float) / 1000.0
for key in ("smooth_block_ms", "hmm_ms"):
    t = np.array([r[key] for r in rows])
    w, f = np.polyfit(c, t, 1)
    print(json.dumps({"stage": key, "fixed_ms": float(f), "ms_per_1000_cells": float(w)}))
