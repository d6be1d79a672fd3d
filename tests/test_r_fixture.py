"""Consumes the output of the REFERENCE's own R functions when a box with R has produced it
(tools/export_r_fixture_inputs.py -> Rscript tools/make_hmm_fixture.R -> ICNV_R_FIXTURE_DIR): Viterbi.dthmm.adj states for
i6 / i3 and .median_filter results on ten c1 cells.  Skipped when the fixture is absent (R is not installable in the build
image); the 50-digit restatement in tests/golden/hmm_mpmath_c1.npz pins the same states meanwhile."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

DIR = os.environ.get("ICNV_R_FIXTURE_DIR", os.path.join(os.path.dirname(__file__), "golden", "r_fixture"))
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(DIR, "out", "i6_states.bin")),
                                reason="no R-generated fixture (run tools/make_hmm_fixture.R on a box with R)")


def _load():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "hmm_mpmath_c1.npz"))
    X = np.asfortranarray(d["X"])
    G, C = X.shape
    rd = lambda f: np.fromfile(os.path.join(DIR, "out", f), dtype="<f8").reshape(C, G).T   # noqa: E731
    return d, X, rd


def test_oracle_states_equal_the_reference_r_output():
    d, X, rd = _load()
    for tag in ("i6", "i3"):
        want = rd(f"{tag}_states.bin")
        got = orc.viterbi_matrix(X, d["chr_start"], d["chr_len"], np.asfortranarray(d[tag + "_Pi"]), d[tag + "_delta"],
                                 d[tag + "_mean"], d[tag + "_sd"])
        np.testing.assert_array_equal(got.astype(float), want, err_msg=tag)
        np.testing.assert_array_equal(d[tag + "_states"].astype(float), want, err_msg=tag + " (50-digit restatement)")


def test_oracle_median_filter_equals_the_reference_r_output():
    d, X, rd = _load()
    lists = [np.arange(0, 4), np.array([4, 6, 5, 8, 7, 9])]
    got = orc.median_filter(X, d["chr_start"], d["chr_len"], lists, 7)
    np.testing.assert_array_equal(got, rd("median_filter.bin"))


@pytest.mark.gpu
def test_gpu_states_and_median_filter_equal_the_reference_r_output():
    from infercnv_b200 import api
    api.init(0)
    d, X, rd = _load()
    for tag in ("i6", "i3"):
        got = api.viterbi(X, d["chr_start"], d["chr_len"], np.asfortranarray(d[tag + "_Pi"]), d[tag + "_delta"], d[tag + "_mean"],
                          d[tag + "_sd"])
        np.testing.assert_array_equal(got.astype(float), rd(f"{tag}_states.bin"), err_msg=tag)
    lists = [np.arange(0, 4), np.array([4, 6, 5, 8, 7, 9])]
    np.testing.assert_array_equal(api.median_filter(X, d["chr_start"], d["chr_len"], lists, 7), rd("median_filter.bin"))
