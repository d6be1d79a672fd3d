#!/usr/bin/env python
"""The reference's bundled example (example/run.R: oligodendroglioma, two reference groups, i6 HMM, denoise) driven
through the Python mirror of the R interface - every numeric step runs in libinfercnv_b200.so on the GPU:

    counts -> gene filters (step 2) -> depth normalisation (3) -> log / subtract / clamp / smooth / centre / subtract /
    2^x (4, 8-12, 14, one fused call) -> i6 HMM on whole samples and per cell (17) -> CNV region reports (17) ->
    proxy expression values (20) -> denoise (22)

What it cannot take from the reference without R: the emission parameters, which run() derives from its RNG-driven
hidden spike (R/inferCNV_HMM.R:15-212); the six (mean, sd) pairs of the reference's bundled `mcmc_obj` fixture are used
instead, with a sqrt(n) trend for the group sizes.

    python examples/run_oligodendroglioma.py [out_dir]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mirror import ops  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    d = np.load(os.path.join(ROOT, "tests", "golden", "oligodendroglioma.npz"))
    fx = np.load(os.path.join(ROOT, "tests", "golden", "hmm_fixture.npz"))
    counts = np.asfortranarray(d["counts"].astype(np.float64))
    G, C = counts.shape
    split = lambda idx, off: [idx[off[i]:off[i + 1]] - 1 for i in range(len(off) - 1)]   # noqa: E731
    refs = dict(zip([str(n) for n in d["ref_names"]], split(d["ref_idx"], d["ref_off"])))
    obs = dict(zip([str(n) for n in d["obs_names"]], split(d["obs_idx"], d["obs_off"])))
    obj = ops.Infercnv(expr_data=counts, count_data=counts.copy(order="F"), gene_order_chr=d["chr_codes"],
                       reference_grouped_cell_indices=refs, observation_grouped_cell_indices=obs,
                       gene_names=["gene_%d" % i for i in range(G)], gene_order_start=np.arange(G) * 1000,
                       gene_order_stop=np.arange(G) * 1000 + 900, cell_names=["cell_%d" % i for i in range(C)],
                       chr_names={i + 1: str(n) for i, n in enumerate(d["chr_levels"])})
    t0 = time.time()
    obj = ops.require_above_min_mean_expr_cutoff(obj, 1.0)                     # cutoff = 1 (example/run.R:15)
    obj = ops.require_above_min_cells_ref(obj, 3)
    obj = ops.normalize_counts_by_seq_depth(obj)
    obj = ops.smooth_block(obj, window_length=101, max_centered_threshold=3.0)
    print("smooth block: %d genes x %d cells, values %.3f .. %.3f" % (*obj.expr_data.shape, obj.expr_data.min(), obj.expr_data.max()))
    cnv_mean_sd = {k: {"mean": m, "sd": s} for k, m, s in zip(ops.CNV_LEVELS, fx["mu"], fx["sd"])}
    fit = {k: (np.log(s), -0.5) for k, s in zip(ops.CNV_LEVELS, fx["sd"])}
    hmm = ops.predict_CNV_via_HMM_on_whole_tumor_samples(obj, True, cnv_mean_sd, fit, t=1e-6)
    ops.generate_cnv_region_reports(hmm, "17_HMM_predHMMi6.hmm_mode-samples", out_dir, ignore_neutral_state=3, by="consensus")
    cells = ops.predict_CNV_via_HMM_on_indiv_cells(obj, cnv_mean_sd, t=1e-6)
    st, n = np.unique(cells.expr_data, return_counts=True)
    print("per-cell i6 states:", dict(zip(st.astype(int).tolist(), n.tolist())))
    proxy = ops.assign_HMM_states_to_proxy_expr_vals(hmm)
    final = ops.clear_noise_via_ref_mean_sd(obj, sd_amplifier=2.0)             # denoise=TRUE, sd_amplifier=2 (run.R:22-23)
    np.save(os.path.join(out_dir, "expr_denoised.npy"), final.expr_data)
    np.save(os.path.join(out_dir, "hmm_proxy_expr.npy"), proxy.expr_data)
    regions = open(os.path.join(out_dir, "17_HMM_predHMMi6.hmm_mode-samples.pred_cnv_regions.dat")).read().splitlines()
    print("%d predicted CNV regions written to %s (%.1f s)" % (len(regions) - 1, out_dir, time.time() - t0))
    for line in regions[:6]:
        print("   ", line)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "examples", "out"))
