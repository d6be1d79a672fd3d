"""Outlier clamp / noise clearing oracle (oracle/denoise.py) against the reference's own known answers
(tests/testthat/test_infer_cnv.R:222-262 clear_noise, :404-433 remove_outliers_norm).  CPU only."""
import numpy as np

from oracle import denoise as ord_

matrix_one = np.arange(1, 6, dtype=float).reshape(5, 1)
matrix_three = np.arange(1, 16, dtype=float).reshape(3, 5).T


def rmat(vals, ncol):
    return np.asarray(vals, dtype=float).reshape(ncol, -1).T       # matrix(vals, ncol=ncol), column-major


def test_clear_noise_known_answers():
    assert np.array_equal(ord_.dot_clear_noise(matrix_one, 0), matrix_one)                          # noise_answer_1
    assert np.array_equal(ord_.dot_clear_noise(matrix_one, 4), rmat([0, 0, 0, 4, 5], 1))             # noise_answer_2
    assert np.array_equal(ord_.dot_clear_noise(matrix_one, 6), np.zeros((5, 1)))                     # noise_answer_3
    assert np.array_equal(ord_.dot_clear_noise(matrix_three, 0), matrix_three)                      # noise_answer_4
    assert np.array_equal(ord_.dot_clear_noise(matrix_three, 12), rmat([0] * 11 + [12, 13, 14, 15], 3))   # noise_answer_5
    assert np.array_equal(ord_.dot_clear_noise(matrix_three, 100), np.zeros((5, 3)))                 # noise_answer_6


def test_remove_outliers_norm_known_answers():
    in_1 = rmat(range(1, 21), 4)
    out_1 = rmat([5] * 5 + list(range(6, 15)) + [15] * 6, 4)
    in_2 = rmat(list(range(1, 16)) + [-5, -4] + list(range(3, 14)) + [21, 26] + list(range(1, 16)) * 2, 4)
    out_2 = rmat(list(range(1, 16)) + [-.5, -.5] + list(range(3, 14)) + [17.75, 17.75] + list(range(1, 16)) * 2, 4)
    assert np.array_equal(ord_.remove_outliers_norm(in_1, lower_bound=-1, upper_bound=30), in_1)
    assert np.array_equal(ord_.remove_outliers_norm(in_1, lower_bound=5, upper_bound=15), out_1)
    assert np.array_equal(ord_.remove_outliers_norm(in_2, out_method="average_bound"), out_2)
    assert ord_.get_average_bounds(in_2) == (-0.5, 17.75)


def test_logistic_adjustment_shape():
    x = np.linspace(0.5, 1.5, 101).reshape(-1, 1)
    y = ord_.apply_logistic_val_adj(x, 1.0, 0.1, 20.0)
    assert y[50, 0] == 1.0                                             # the centre is left alone
    assert np.all(np.abs(y - 1.0) <= np.abs(x - 1.0) + 1e-16)          # values are pulled towards the centre
    assert abs(y[0, 0] - 0.5) < 1e-3 and abs(y[49, 0] - 1.0) < 0.2 * 0.01    # far values kept, near ones flattened
    assert np.allclose(y - 1.0, -(y[::-1] - 1.0), atol=1e-14)          # odd symmetry about the centre
