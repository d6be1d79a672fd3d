"""predict_CNV_via_HMM_on_tumor_subclusters_per_chr (R/inferCNV_HMM.R:412-487): a different partition of the cells on
every chromosome, one Viterbi trace per (chromosome, subcluster) on its rowMeans, then the per-subcluster consensus.
Against the oracle's Viterbi run chromosome by chromosome on the row block, and its consensus restatement."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import regions as orr

pytestmark = pytest.mark.gpu


def _setup(example_object, hmm_fixture):
    from mirror import ops
    ex = example_object
    X = orc.smooth_block(orc.normalize_by_seq_depth(ex["counts"]), *orc.chr_ranges(ex["chr_codes"]), ex["ref_groups"])
    obj = ops.Infercnv(expr_data=X, gene_order_chr=ex["chr_codes"],
                       reference_grouped_cell_indices={"normal": ex["ref_groups"][0]},
                       observation_grouped_cell_indices={"tumor": ex["obs_groups"][0]},
                       tumor_subclusters={"subclusters": {"tumor": {"tumor_s1": ex["subclusters"][0][:6],
                                                                    "tumor_s2": ex["subclusters"][0][6:]}}})
    cnv_mean_sd = {k: {"mean": m, "sd": s} for k, m, s in zip(ops.CNV_LEVELS, hmm_fixture["mean"], hmm_fixture["sd"])}
    fit = {k: (np.log(s), -0.5) for k, s in zip(ops.CNV_LEVELS, hmm_fixture["sd"])}
    return obj, cnv_mean_sd, fit


def test_per_chromosome_subcluster_hmm_and_consensus(example_object, hmm_fixture):
    from infercnv_b200 import api
    from mirror import ops
    obj, cnv_mean_sd, fit = _setup(example_object, hmm_fixture)
    X = obj.expr_data
    G, C = X.shape
    cs, cl = orc.chr_ranges(example_object["chr_codes"])
    rng = np.random.default_rng(1)
    obs = example_object["obs_groups"][0]
    per_chr = {}
    for k, s in enumerate(cs):            # a fresh random partition of the observation cells on every chromosome
        perm = rng.permutation(obs)
        cut = sorted(rng.choice(np.arange(1, len(perm)), size=k % 3, replace=False).tolist())
        per_chr[example_object["chr_codes"][s]] = [g for g in np.split(perm, cut)]
    Pi, delta = orc.hmm_params(6)
    # oracle: per chromosome, Viterbi of the row block with that chromosome's groups
    want = np.full((G, C), 255, dtype=np.uint8)
    for k, (s, n) in enumerate(zip(cs, cl)):
        groups = per_chr[example_object["chr_codes"][s]]
        sds = np.concatenate([hmm_fixture["sd"] * len(g) ** -0.5 for g in groups])
        blk = orc.viterbi_matrix(np.asfortranarray(X[s:s + n]), [0], [n], Pi, delta, hmm_fixture["mean"], sds, groups=groups)
        want[s:s + n] = np.where(blk < 0, 255, blk).astype(np.uint8)
    per = [per_chr[example_object["chr_codes"][s]] for s in cs]
    sds_all = np.concatenate([hmm_fixture["sd"] * len(g) ** -0.5 for p in per for g in p])
    got = api.viterbi_per_chr(X, cs, cl, per, Pi, delta, hmm_fixture["mean"], sds_all)
    assert np.array_equal(got, want)
    ref_cells = example_object["ref_groups"][0]
    assert np.all(got[:, ref_cells] == 255)                 # references are in no per-chromosome subcluster here
    # consensus broadcast (HMM.R:470-483)
    subs = [np.asarray(v) for v in obj.tumor_subclusters["subclusters"]["tumor"].values()]
    want_c = want.copy()
    valid = np.repeat(np.asarray(cl) >= 2, cl)
    for g in subs:
        cons = orr.state_consensus(want, g)
        want_c[np.ix_(valid, g)] = cons[valid][:, None]
    assert np.array_equal(api.apply_state_consensus(got, cs, cl, subs), want_c)
    out = ops.predict_CNV_via_HMM_on_tumor_subclusters_per_chr(obj, per_chr, cnv_mean_sd, fit, t=1e-6)
    assert np.array_equal(out.expr_data, np.where(want_c == 255, -1.0, want_c.astype(float)))
    # a cell listed twice on one chromosome is rejected
    bad = [list(p) for p in per]
    bad[0] = bad[0] + [bad[0][0][:1]]
    with pytest.raises(Exception):
        api.viterbi_per_chr(X, cs, cl, bad, Pi, delta, hmm_fixture["mean"], np.tile(hmm_fixture["sd"], sum(len(p) for p in bad)))
