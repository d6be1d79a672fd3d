"""CPU oracle for the outlier clamp and noise clearing steps (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

* ``.remove_outliers_norm`` / ``.get_average_bounds``  R/inferCNV_ops.R:1998-2056, 2734-2742
* ``clear_noise`` / ``.clear_noise``                   R/inferCNV_ops.R:2232-2275
* ``.apply_logistic_val_adj`` / ``.logistic``          R/inferCNV_heatmap.R:2792-2810, R/SplatterScrape.R:210-212
* ``clear_noise_via_ref_mean_sd(noise_logistic=TRUE)`` R/inferCNV_ops.R:2325-2329

Pinned by the reference's own known answers (tests/testthat/test_infer_cnv.R:222-262 and 404-433).
"""
from __future__ import annotations

import numpy as np


def r_mean(x) -> float:
    """R's mean(): long-double sum and one refinement pass (summary.c)."""
    x = np.asarray(x, dtype=np.float64).ravel(order="F")
    s = np.longdouble(0)
    for v in x:
        s += v
    s /= len(x)
    t = np.longdouble(0)
    for v in x:
        t += v - s
    return float(s + t / len(x))


def get_average_bounds(expr_matrix):
    X = np.asarray(expr_matrix, dtype=np.float64)
    return r_mean(np.nanmin(X, axis=0)), r_mean(np.nanmax(X, axis=0))       # quantile(x, na.rm=TRUE)[[1]] / [[5]]


def remove_outliers_norm(data, out_method="average_bound", lower_bound=None, upper_bound=None):
    X = np.array(data, dtype=np.float64, order="F")
    if lower_bound is None or upper_bound is None:
        if out_method != "average_bound":
            raise RuntimeError("991")
        lower_bound, upper_bound = get_average_bounds(X)
    X[X < lower_bound] = lower_bound
    X[X > upper_bound] = upper_bound
    return X


def dot_clear_noise(expr_data, threshold, center_pos=0.0):
    X = np.array(expr_data, dtype=np.float64, order="F")
    X[(X > center_pos - threshold) & (X < center_pos + threshold)] = center_pos
    return X


def apply_logistic_val_adj(vals, expr_mean, delta_midpt, slope=20.0):
    X = np.asarray(vals, dtype=np.float64)
    val = np.abs(X - expr_mean)
    p = 1.0 / (1.0 + np.exp(-slope * (val - delta_midpt)))
    return np.asfortranarray(np.where(X > expr_mean, expr_mean + p * val, np.where(X < expr_mean, expr_mean - p * val, X)))


def clear_noise(expr_data, ref_idx, threshold, noise_logistic=False):
    X = np.asarray(expr_data, dtype=np.float64)
    if threshold == 0:
        return np.asfortranarray(X.copy())
    centre = r_mean(X[:, np.asarray(ref_idx)] if ref_idx is not None and len(ref_idx) else X)
    return apply_logistic_val_adj(X, centre, threshold) if noise_logistic else dot_clear_noise(X, threshold, centre)


def clear_noise_via_ref_mean_sd_logistic(expr_data, ref_idx, sd_amplifier=1.5):
    X = np.asarray(expr_data, dtype=np.float64)
    vals = X[:, np.asarray(ref_idx)]
    mean_ref_sd = r_mean(np.std(vals, axis=0, ddof=1)) * sd_amplifier
    return apply_logistic_val_adj(X, r_mean(vals), mean_ref_sd)
