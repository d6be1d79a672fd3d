"""The stand-alone element-wise / reduction entry points of icnv_reduce.cu through the C ABI against the oracle: the
steps either side of the smooth block when run() is driven step by step.  Also executed by the host emulation
(tests/test_emulated_kernels.py), like the other test_gpu_widen_* files."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_elementwise_steps(example_object):
    from infercnv_b200 import api
    X = orc.normalize_by_seq_depth(example_object["counts"])
    L = api.log2xplus1(X)
    np.testing.assert_allclose(L, orc.log2xplus1(X), rtol=1e-15, atol=1e-16)
    T = api.apply_max_threshold_bounds(L - 2.0, 1.5)
    assert np.array_equal(T, orc.apply_max_threshold_bounds(L - 2.0, 1.5))
    np.testing.assert_allclose(api.invert_log2(T), orc.invert_log2(T), rtol=1e-15)
    with pytest.raises(Exception):
        bad = X.copy(order="F")
        bad[3, 2] = np.nan
        api.log2xplus1(bad)                                   # non-finite input is an error, as in the fused block


def test_depth_normalisation_and_denoise(example_object):
    from infercnv_b200 import api
    C = example_object["counts"]
    for nf in (None, 1e5):
        assert np.array_equal(api.normalize_counts_by_seq_depth(C, nf), orc.normalize_by_seq_depth(C, nf))
    E = example_object["expr"]
    ref = np.concatenate(example_object["ref_groups"])
    mu, sg = api.mean_sd(E, ref)
    mu_o, sg_o = orc.mean_sd_over_cells(E, ref)
    assert abs(mu - mu_o) <= 1e-14 * abs(mu_o) and abs(sg - sg_o) <= 1e-13 * sg_o
    got = api.clear_noise_via_ref_mean_sd(E, ref, 2.0)
    want = orc.clear_noise_via_ref_mean_sd(E, ref, 2.0)
    assert np.mean(np.abs(got - want) > 1e-14 * np.abs(want)) < 1e-6


def test_proxy_expression_values():
    from infercnv_b200 import api
    rng = np.random.default_rng(0)
    S = rng.integers(1, 7, size=(500, 33)).astype(np.float64)
    S[rng.random(S.shape) < 0.05] = -1.0
    lut6 = np.array([-1.0, 0.0, 0.5, 1.0, 1.5, 2.0, 3.0])
    want = np.where(S < 0, -1.0, lut6[np.maximum(S, 0).astype(int)])
    assert np.array_equal(api.assign_hmm_states_to_proxy_expr_vals(S, 6), want)
    S3 = np.clip(S, -1, 3)
    lut3 = np.array([-1.0, 0.5, 1.0, 1.5])
    assert np.array_equal(api.assign_hmm_states_to_proxy_expr_vals(S3, 3), np.where(S3 < 0, -1.0, lut3[np.maximum(S3, 0).astype(int)]))
