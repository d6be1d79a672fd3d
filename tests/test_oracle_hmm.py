"""Oracle self-consistency for the parts the reference has no fixture for (Viterbi, median filter):
C restatement vs an independent pure-Python transcription, plus the documented quirks."""
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _rand_seq(rng, n, mean):
    lv = rng.integers(0, len(mean), size=max(1, n // 40 + 1))
    x = np.repeat(np.asarray(mean)[lv], 40)[:n] + rng.normal(scale=0.08, size=n)
    return x


@pytest.mark.parametrize("m", [6, 3])
def test_viterbi_c_vs_python_literal(hmm_fixture, m):
    rng = np.random.default_rng(11 + m)
    if m == 6:
        mean, sd = hmm_fixture["mean"], hmm_fixture["sd"]
    else:
        mean, sd = np.array([0.9, 1.0, 1.1]), np.array([0.05] * 3)
    Pi, delta = orc.hmm_params(m)
    worst = np.inf
    for n in [2, 3, 17, 90, 400, 852]:
        for _ in range(4):
            x = _rand_seq(rng, n, mean)
            st, margin = orc.viterbi_seq(x, Pi, delta, mean, sd)
            py = orc.literal_viterbi(x, Pi, delta, mean, sd)
            if margin > 1e-9:          # outside an ulp-level tie both must agree exactly
                np.testing.assert_array_equal(st, py)
            worst = min(worst, margin)
            assert st.min() >= 1 and st.max() <= m
    assert worst > 0


def test_viterbi_quirks(hmm_fixture):
    mean, sd = hmm_fixture["mean"], hmm_fixture["sd"]
    Pi6, d6 = orc.hmm_params(6)
    Pi3, d3 = orc.hmm_params(3)
    # Q5: diagonal 1-5t for both models; i3 rows sum to 1-3t
    assert np.isclose(Pi3.sum(axis=1), 1 - 3e-6).all() and np.isclose(Pi6.sum(axis=1), 1.0).all()
    assert d6[2] == 1 - 5e-6 and d3[1] == 1 - 5e-6
    # Q8: a chromosome with < 2 genes gets state 3 regardless of the model (HMM.R:1104-1107)
    st, _ = orc.viterbi_seq(np.array([1.0]), Pi3, d3, np.array([0.9, 1.0, 1.1]), np.array([0.05] * 3))
    assert st.tolist() == [3]
    # neutral input stays neutral; a long amplified stretch is called
    x = np.full(300, mean[2])
    st, _ = orc.viterbi_seq(x, Pi6, d6, mean, sd)
    assert (st == 3).all()
    x[100:200] = mean[4]
    st, _ = orc.viterbi_seq(x, Pi6, d6, mean, sd)
    assert (st[:100] == 3).all() and (st[100:200] == 5).all() and (st[200:] == 3).all()
    # non-finite input -> error, as the reference would stop (HMM.R:1165)
    with pytest.raises(ValueError):
        orc.viterbi_seq(np.array([1.0, np.nan, 1.0]), Pi6, d6, mean, sd)


def test_viterbi_matrix_modes(hmm_fixture):
    rng = np.random.default_rng(5)
    mean, sd = hmm_fixture["mean"], hmm_fixture["sd"]
    Pi, delta = orc.hmm_params(6)
    G, C = 300, 12
    cs, cl = np.array([0, 120, 121]), np.array([120, 1, 179])
    X = np.asfortranarray(np.column_stack([_rand_seq(rng, G, mean) for _ in range(C)]))
    st, mg = orc.viterbi_matrix(X, cs, cl, Pi, delta, mean, sd, want_margins=True, nthreads=2)
    for c in range(C):
        for k in range(3):
            s1, _ = orc.viterbi_seq(X[cs[k]:cs[k] + cl[k], c], Pi, delta, mean, sd)
            np.testing.assert_array_equal(st[cs[k]:cs[k] + cl[k], c], s1)
    assert (st[120] == 3).all()
    # group mode: rowMeans over the group, trace broadcast, outsiders stay -1 (HMM.R:368,383,399)
    groups = [np.array([0, 3, 5]), np.array([7, 8])]
    sg = orc.viterbi_matrix(X, cs, cl, Pi, delta, mean, np.tile(sd, 2), groups=groups)
    for g in groups:
        xm = X[:, g].mean(axis=1)
        for k in range(3):
            s1, _ = orc.viterbi_seq(xm[cs[k]:cs[k] + cl[k]], Pi, delta, mean, sd)
            for c in g:
                np.testing.assert_array_equal(sg[cs[k]:cs[k] + cl[k], c], s1)
    outsiders = sorted(set(range(C)) - {0, 3, 5, 7, 8})
    assert (sg[:, outsiders] == -1).all()


def test_median_filter_vs_numpy():
    rng = np.random.default_rng(3)
    G, C = 60, 25
    X = np.asfortranarray(rng.normal(size=(G, C)))
    cs, cl = np.array([0, 37]), np.array([37, 23])
    groups = [np.array([4, 2, 9, 11, 0, 1, 3, 5, 6, 7, 8, 10]), np.array([20, 21, 12]), np.arange(13, 20)]
    got = orc.median_filter(X, cs, cl, groups, window_size=7, nthreads=2)
    want = X.copy()
    r = 4  # (window_size+1)/2: noise_reduction.R:102-106 uses half_window+1
    for g in groups:
        for s, n in zip(cs, cl):
            B = X[s:s + n][:, g]
            out = np.empty_like(B)
            for i in range(n):
                for j in range(len(g)):
                    out[i, j] = np.median(B[max(0, i - r):min(n, i + r + 1), max(0, j - r):min(len(g), j + r + 1)])
            want[s:s + n, g] = out
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-15)
    # cells in no group are untouched
    np.testing.assert_array_equal(got[:, [22, 23, 24]], X[:, [22, 23, 24]])


def test_i3_params(example_object):
    ex = example_object
    X = orc.normalize_by_seq_depth(ex["counts"])
    ref = np.concatenate(ex["ref_groups"])
    mu, sg = orc.mean_sd_over_cells(X, ref)
    vals = X[:, ref].ravel(order="F")
    assert np.isclose(mu, vals.mean(), rtol=1e-13) and np.isclose(sg, vals.std(ddof=1), rtol=1e-12)
    Pi, delta, mean, sd = orc.i3_hmm_params(X, ref)
    assert mean[0] < mean[1] < mean[2] and np.isclose(mean[2] - mean[1], 1.6448536269514722 * sg)


def test_oracle_viterbi_equals_the_50_digit_restatement_on_c1():
    """tests/golden/hmm_mpmath_c1.npz (tools/make_hmm_mpmath_fixture.py): Viterbi.dthmm.adj evaluated with 50 significant
    digits on ten cells of the bundled oligodendroglioma example (all 22 chromosomes, i6 and i3), with the smallest arg-max
    margin of every sequence.  The margins (>= 2e-5) are seven orders of magnitude above what double-precision rounding can
    accumulate over a chromosome, so the exact states are what any faithful double-precision evaluation - the reference's R
    included - must return; the C oracle returns them."""
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "hmm_mpmath_c1.npz"))
    X = np.asfortranarray(d["X"])
    for tag in ("i6", "i3"):
        assert d[tag + "_margins"].min() > 1e-6
        got = orc.viterbi_matrix(X, d["chr_start"], d["chr_len"], np.asfortranarray(d[tag + "_Pi"]), d[tag + "_delta"],
                                 d[tag + "_mean"], d[tag + "_sd"], nthreads=orc.max_threads())
        np.testing.assert_array_equal(got, d[tag + "_states"].astype(got.dtype), err_msg=tag)
