"""run() step 16 (remove_outliers_norm) and step 22 with a numeric noise_filter / noise_logistic (clear_noise,
depress_log_signal_midpt_val) through the C ABI against the reference's known answers
(tests/testthat/test_infer_cnv.R:222-262, 404-433) and the oracle."""
import numpy as np
import pytest

from oracle import denoise as ord_

pytestmark = pytest.mark.gpu


def rmat(vals, ncol):
    return np.asfortranarray(np.asarray(vals, dtype=float).reshape(ncol, -1).T)


def test_remove_outliers_norm_known_answers():
    from infercnv_b200 import api
    in_1 = rmat(range(1, 21), 4)
    out_1 = rmat([5] * 5 + list(range(6, 15)) + [15] * 6, 4)
    in_2 = rmat(list(range(1, 16)) + [-5, -4] + list(range(3, 14)) + [21, 26] + list(range(1, 16)) * 2, 4)
    out_2 = rmat(list(range(1, 16)) + [-.5, -.5] + list(range(3, 14)) + [17.75, 17.75] + list(range(1, 16)) * 2, 4)
    assert np.array_equal(api.remove_outliers_norm(in_1, -1, 30), in_1)
    assert np.array_equal(api.remove_outliers_norm(in_1, 5, 15), out_1)
    got, bounds = api.remove_outliers_norm(in_2, want_bounds=True)
    assert np.array_equal(got, out_2) and bounds == (-0.5, 17.75)


def test_remove_outliers_norm_mirror_on_the_reference_example(example_object):
    from infercnv_b200 import api
    from mirror import ops
    X = example_object["expr"]
    H = np.asfortranarray(X[:, :7] * 1.5)
    obj = ops.Infercnv(expr_data=X, gene_order_chr=example_object["chr_codes"],
                       hspike=ops.Infercnv(expr_data=H, gene_order_chr=example_object["chr_codes"]))
    out = ops.remove_outliers_norm(obj)                                   # average_bound
    assert np.array_equal(out.expr_data, ord_.remove_outliers_norm(X))
    assert np.array_equal(out.hspike.expr_data, ord_.remove_outliers_norm(H))      # mirrored onto @.hspike
    assert np.array_equal(ops.remove_outliers_norm(obj, None, 0.9, 1.1).expr_data, np.clip(X, 0.9, 1.1))
    with pytest.raises(RuntimeError):
        ops.remove_outliers_norm(obj, out_method="quantile")              # stop(991)
    with pytest.raises(RuntimeError):
        ops.remove_outliers_norm(obj, out_method=None)                    # stop(992)
    Xn = X.copy(order="F")
    Xn[5, 3] = np.nan                                                     # quantile(na.rm=TRUE): NaN is skipped
    got, b = api.remove_outliers_norm(Xn, want_bounds=True)
    assert b == ord_.get_average_bounds(Xn) and np.isnan(got[5, 3])


def test_clear_noise_known_answers_and_modes(example_object):
    from infercnv_b200 import api
    from mirror import ops
    m3 = rmat(range(1, 16), 3)
    # .clear_noise(expr, threshold) with center_pos = 0 is what the reference's unit tests pin; through clear_noise()
    # the centre is the mean of the reference cells, so shift the data to put that mean at 0
    ref = np.array([1])
    centre = m3[:, 1].mean()
    got = api.clear_noise(m3 - centre, ref, 3)
    assert np.array_equal(got + centre, ord_.dot_clear_noise(m3, 3, centre))
    X = example_object["expr"]
    refs = np.concatenate(example_object["ref_groups"])
    obj = ops.Infercnv(expr_data=X, gene_order_chr=example_object["chr_codes"],
                       reference_grouped_cell_indices={"normal": refs},
                       observation_grouped_cell_indices={"tumor": np.concatenate(example_object["obs_groups"])})
    assert ops.clear_noise(obj, 0) is obj                                  # threshold 0: nothing to do
    for thr in (0.02, 0.1):
        want = ord_.clear_noise(X, refs, thr)
        got = ops.clear_noise(obj, thr).expr_data
        # the centre is a mean over 46 130 values: the library's summation order differs from R's long-double
        # accumulation by an ulp, so cleared values agree to rounding, and a value within an ulp of a bound may flip
        assert np.mean(np.abs(got - want) > 1e-14 * np.abs(want)) < 1e-6
        assert np.allclose(got, want, rtol=0, atol=thr * 1.0001)
        assert 0 < (got != X).sum() < X.size
        gl = ops.clear_noise(obj, thr, noise_logistic=True).expr_data
        assert np.allclose(gl, ord_.clear_noise(X, refs, thr, noise_logistic=True), rtol=1e-12, atol=1e-14)
    obj_noref = ops.Infercnv(expr_data=X, gene_order_chr=example_object["chr_codes"],
                             observation_grouped_cell_indices={"all": np.arange(X.shape[1])})
    want = ord_.clear_noise(X, None, 0.05)
    assert np.mean(np.abs(ops.clear_noise(obj_noref, 0.05).expr_data - want) > 1e-14 * np.abs(want)) < 1e-6
    gl = ops.clear_noise_via_ref_mean_sd(obj, 1.5, noise_logistic=True).expr_data
    assert np.allclose(gl, ord_.clear_noise_via_ref_mean_sd_logistic(X, refs, 1.5), rtol=1e-12, atol=1e-14)
