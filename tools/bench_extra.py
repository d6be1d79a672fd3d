#!/usr/bin/env python
"""Timings of the secondary kernels at configs[1] size (10 000 x 10 000): 2-D median filter
(apply_median_filtering, window 7, subclusters of 50-500 cells) and the i3 HMM.  CUDA events, 5 runs."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from infercnv_b200 import dist as shard  # noqa: E402
from infercnv_b200.device import Engine  # noqa: E402
from infercnv_b200.ops import i3HMM_get_HMM  # noqa: E402

eng = Engine(0)
G, C = 10000, 10000
cs, cl = bench.chr_layout(G)
refs = bench.ref_groups_global(C)
plan = shard.plan_shards(C, refs, 1)[0]
X = eng.synth(G, cs, cl, plan.local_cells, C, bench.SEED)
Y, _ = eng.smooth_block(X, cs, cl, plan.local_ref_groups(), plan.ref_sizes, plan.max_chunks)
rng = np.random.default_rng(4)
groups, pos = [], 0
while pos < C:
    n = int(rng.integers(50, 501))
    groups.append(np.arange(pos, min(C, pos + n)))
    pos += n


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
F = torch.empty_like(Y)
ms = timed(lambda: eng.median_filter(Y, cs, cl, groups, 7, out=F))
out["median_filter_w7"] = {"ms": ms, "cell_genes_per_s": G * C / ms * 1e3, "algorithmic_GBps": 16 * G * C / ms / 1e6,
                           "blocks": len(groups) * len(cs)}
ref_cells = np.concatenate(plan.local_ref_groups())
mu = float(Y[ref_cells].mean().item())
sg = float(Y[ref_cells].std().item())
from statistics import NormalDist  # noqa: E402
Pi3, d3, mean3, sd3 = i3HMM_get_HMM({"mu": mu, "sigma": sg, "mean_delta": abs(NormalDist(0, sg).inv_cdf(0.05)), "KS_delta": None}, 1e-6)
ms = timed(lambda: eng.viterbi(Y, cs, cl, Pi3, d3, mean3, sd3))
out["viterbi_i3"] = {"ms": ms, "cell_genes_per_s": G * C / ms * 1e3, "algorithmic_GBps": 9 * G * C / ms / 1e6}
print(json.dumps(out))
