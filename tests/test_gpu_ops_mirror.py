"""The host-side mirror of the R interface (mirror/ops.py, test infrastructure): the reference's own step sequence
(run() steps 4, 8, 9, 10, 11, 12, 14 - R/inferCNV_ops.R:614-1031) driven function by function, as
example/example.Rmd and run(up_to_step=...) do, against the reference's bundled result."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _obj(ex, with_hspike=False):
    from mirror import ops
    X = orc.normalize_by_seq_depth(ex["counts"])           # step 3 (out of scope, test infrastructure)
    o = ops.Infercnv(expr_data=X, gene_order_chr=ex["chr_codes"],
                     reference_grouped_cell_indices={"normal": ex["ref_groups"][0]},
                     observation_grouped_cell_indices={"tumor": ex["obs_groups"][0]},
                     tumor_subclusters={"subclusters": {"tumor": {"tumor_s1": ex["subclusters"][0]},
                                                        "normal": {"normal_s1": ex["subclusters"][1]}}})
    if with_hspike:
        rng = np.random.default_rng(0)
        H = rng.poisson(3.0, size=(X.shape[0], 12)).astype(float)
        o.hspike = ops.Infercnv(expr_data=H, gene_order_chr=ex["chr_codes"],
                                reference_grouped_cell_indices={"simnormal": np.arange(0, 6)},
                                observation_grouped_cell_indices={"simtumor": np.arange(6, 12)})
    return o


def test_stepwise_run_sequence_reproduces_the_reference_golden(example_object):
    from infercnv_b200 import api
    from mirror import ops
    api.init(0)
    ex = example_object
    o = _obj(ex, with_hspike=True)
    o = ops.log2xplus1(o)                                               # step 4
    o = ops.subtract_ref_expr_from_obs(o, inv_log=False, use_bounds=True)   # step 8
    o = ops.apply_max_threshold_bounds(o, threshold=3)                  # step 9
    o = ops.smooth_by_chromosome(o, window_length=101, smooth_ends=True)    # step 10
    o = ops.center_cell_expr_across_chromosome(o, method="median")      # step 11
    o = ops.subtract_ref_expr_from_obs(o, inv_log=False, use_bounds=True)   # step 12
    o = ops.invert_log2(o)                                              # step 14
    ref = np.concatenate(ex["ref_groups"])
    final = orc.clear_noise_via_ref_mean_sd(o.expr_data, ref, 1.5)      # step 22 (denoise) - oracle, not product
    rel = np.max(np.abs(final - ex["expr"]) / np.abs(ex["expr"]))
    print(f"\n[ops mirror, stepwise] max rel err vs the reference's expr.data: {rel:.3e}")
    assert rel < 1e-5 and rel < 1e-11
    # fused block == stepwise (to rounding), on the main matrix and on the mirrored hspike
    f = ops.smooth_block(_obj(ex, with_hspike=True))
    assert np.max(np.abs(f.expr_data - o.expr_data) / np.abs(o.expr_data)) < 1e-12
    assert o.hspike is not None and f.hspike is not None
    assert np.max(np.abs(f.hspike.expr_data - o.hspike.expr_data) / np.abs(o.hspike.expr_data)) < 1e-12
    # hspike mirroring really ran the same step on the spike matrix
    want_h = orc.smooth_block(_obj(ex, with_hspike=True).hspike.expr_data, *orc.chr_ranges(ex["chr_codes"]),
                              [np.arange(0, 6)])
    assert np.max(np.abs(o.hspike.expr_data - want_h) / np.abs(want_h)) < 1e-11


def test_whole_bundled_example_runs_on_the_gpu_counts_to_expr(example_object):
    """count.data -> expr.data of data/infercnv_object_example.rda with EVERY numeric step in the library
    (steps 3, 4..14 fused, 22): nothing of the oracle on the path, only the reference's stored answer."""
    from mirror import ops
    ex = example_object
    o = ops.Infercnv(expr_data=ex["counts"], gene_order_chr=ex["chr_codes"],
                     reference_grouped_cell_indices={"normal": ex["ref_groups"][0]},
                     observation_grouped_cell_indices={"tumor": ex["obs_groups"][0]})
    o = ops.normalize_counts_by_seq_depth(o)
    o = ops.smooth_block(o, window_length=101, max_centered_threshold=3.0)
    o = ops.clear_noise_via_ref_mean_sd(o, sd_amplifier=1.5)
    rel = np.max(np.abs(o.expr_data - ex["expr"]) / np.abs(ex["expr"]))
    print(f"\n[bundled example, all steps on the GPU] max rel err vs the reference's expr.data: {rel:.3e}")
    assert rel < 1e-5 and rel < 1e-11
    from infercnv_b200 import api
    np.testing.assert_allclose(api.normalize_counts_by_seq_depth(ex["counts"]), orc.normalize_by_seq_depth(ex["counts"]),
                               rtol=1e-13)
    np.testing.assert_allclose(api.normalize_counts_by_seq_depth(ex["counts"], 1e4).sum(axis=0), 1e4, rtol=1e-12)
    ref = np.concatenate(ex["ref_groups"])
    np.testing.assert_allclose(api.clear_noise_via_ref_mean_sd(ex["expr"], ref, 2.0),
                               orc.clear_noise_via_ref_mean_sd(ex["expr"], ref, 2.0), rtol=1e-12)


def test_hmm_drivers_and_median_filter_through_the_mirror(example_object, hmm_fixture):
    from mirror import ops
    ex = example_object
    o = ops.smooth_block(_obj(ex))
    cnv_mean_sd = {k: {"mean": m, "sd": s} for k, m, s in zip(ops.CNV_LEVELS, hmm_fixture["mean"], hmm_fixture["sd"])}
    cs, cl = orc.chr_ranges(ex["chr_codes"])
    Pi, delta = orc.hmm_params(6)
    cells = ops.predict_CNV_via_HMM_on_indiv_cells(o, cnv_mean_sd, t=1e-6)
    want = orc.viterbi_matrix(o.expr_data, cs, cl, Pi, delta, hmm_fixture["mean"], hmm_fixture["sd"])
    np.testing.assert_array_equal(cells.expr_data, want.astype(float))
    # group modes: per-group sds from a log-log trend as .get_state_emission_params computes them (HMM.R:586-614)
    fit = {k: (np.log(s), -0.5) for k, s in zip(ops.CNV_LEVELS, hmm_fixture["sd"])}
    samples = ops.predict_CNV_via_HMM_on_whole_tumor_samples(o, True, cnv_mean_sd, fit, t=1e-6)
    groups = [ex["obs_groups"][0], ex["ref_groups"][0]]
    sds = np.concatenate([hmm_fixture["sd"] * len(g) ** -0.5 for g in groups])
    want_g = orc.viterbi_matrix(o.expr_data, cs, cl, Pi, delta, hmm_fixture["mean"], sds, groups=groups)
    np.testing.assert_array_equal(samples.expr_data, want_g.astype(float))
    # cluster_by_groups = FALSE: c(all_observations = unlist(obs), reference_list) in the reference (HMM.R:531) makes every
    # observation cell its own one-cell sample (sd for num_cells = 1), the reference groups stay groups
    flat = ops.predict_CNV_via_HMM_on_whole_tumor_samples(o, False, cnv_mean_sd, fit, t=1e-6)
    groups1 = [np.asarray([c]) for c in ex["obs_groups"][0]] + [ex["ref_groups"][0]]
    sds1 = np.concatenate([hmm_fixture["sd"] * len(g) ** -0.5 for g in groups1])
    want_1 = orc.viterbi_matrix(o.expr_data, cs, cl, Pi, delta, hmm_fixture["mean"], sds1, groups=groups1)
    np.testing.assert_array_equal(flat.expr_data, want_1.astype(float))
    obs_cols = ex["obs_groups"][0]   # a one-cell sample at sd(num_cells = 1) is the per-cell HMM of that cell
    np.testing.assert_array_equal(flat.expr_data[:, obs_cols], cells.expr_data[:, obs_cols])
    i3flat = ops.i3HMM_predict_CNV_via_HMM_on_whole_tumor_samples(o, False, i3_p_val=0.05, t=1e-6, use_KS=False)
    Pi3f, d3f, mean3f, sd3f = orc.i3_hmm_params(o.expr_data, ex["ref_groups"][0])
    want_3f = orc.viterbi_matrix(o.expr_data, cs, cl, Pi3f, d3f, mean3f, np.tile(sd3f, len(groups1)), groups=groups1)
    np.testing.assert_array_equal(i3flat.expr_data, want_3f.astype(float))
    sub = ops.predict_CNV_via_HMM_on_tumor_subclusters(o, cnv_mean_sd, fit, t=1e-6)
    sgroups = [ex["subclusters"][0], ex["subclusters"][1]]
    want_s = orc.viterbi_matrix(o.expr_data, cs, cl, Pi, delta, hmm_fixture["mean"], sds, groups=sgroups)
    np.testing.assert_array_equal(sub.expr_data, want_s.astype(float))
    # i3 per cell with mu / sigma from the reference cells
    i3 = ops.i3HMM_predict_CNV_via_HMM_on_indiv_cells(o, i3_p_val=0.05, t=1e-6, use_KS=False)
    Pi3, d3, mean3, sd3 = orc.i3_hmm_params(o.expr_data, ex["ref_groups"][0])
    want3 = orc.viterbi_matrix(o.expr_data, cs, cl, Pi3, d3, mean3, sd3)
    np.testing.assert_array_equal(i3.expr_data, want3.astype(float))
    assert set(np.unique(i3.expr_data)).issubset({1.0, 2.0, 3.0})
    # proxy values (HMM.R:1191-1206)
    proxy = ops.assign_HMM_states_to_proxy_expr_vals(cells)
    assert set(np.unique(proxy.expr_data)).issubset({0.0, 0.5, 1.0, 1.5, 2.0, 3.0})
    lut6 = np.array([np.nan, 0.0, 0.5, 1.0, 1.5, 2.0, 3.0])
    np.testing.assert_array_equal(proxy.expr_data, lut6[cells.expr_data.astype(int)])
    proxy3 = ops.i3HMM_assign_HMM_states_to_proxy_expr_vals(i3)                       # i3HMM.R:405-417
    np.testing.assert_array_equal(proxy3.expr_data, np.array([np.nan, 0.5, 1.0, 1.5])[i3.expr_data.astype(int)])
    tiny = ops.Infercnv(expr_data=np.array([[-1.0, 1.0, 4.0], [6.0, 3.0, 2.5]]), gene_order_chr=np.array([1, 1]))
    np.testing.assert_array_equal(ops.assign_HMM_states_to_proxy_expr_vals(tiny).expr_data, [[-1.0, 0.0, 1.5], [3.0, 1.0, 2.5]])
    # exported apply_median_filtering: subclusters in hclust order for observations, whole groups for references
    mf = ops.apply_median_filtering(o, window_size=7)
    want_mf = orc.median_filter(o.expr_data, cs, cl, [ex["subclusters"][0], ex["ref_groups"][0]], 7)
    np.testing.assert_allclose(mf.expr_data, want_mf, rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        ops.apply_median_filtering(o, window_size=4)


def test_parallelDist_as_the_reference_calls_it_before_hclust(example_object):
    """hc <- hclust(parallelDist(t(tumor_expr_data), threads = ...)) - R/inferCNV_tumor_subclusters.R:191: the distance vector
    of the observation cells of the bundled example, against the oracle and (when scipy is there) scipy's own pdist."""
    from mirror import ops
    expr = example_object["expr"]
    obs = np.concatenate(example_object["obs_groups"])
    d = ops.parallelDist(expr[:, obs].T, threads=4)
    np.testing.assert_allclose(d, orc.pairwise_dist(expr, obs), rtol=1e-13, atol=0)
    try:
        from scipy.spatial.distance import pdist
        from scipy.cluster.hierarchy import linkage
    except ImportError:
        return
    np.testing.assert_allclose(d, pdist(expr[:, obs].T), rtol=1e-13, atol=0)
    # the tree hclust(method = "ward.D2") builds from it is the tree built from scipy's distances (merge heights equal)
    np.testing.assert_allclose(linkage(d, method="ward")[:, 2], linkage(pdist(expr[:, obs].T), method="ward")[:, 2], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        ops.parallelDist(expr[:, obs].T, method="manhattan")
