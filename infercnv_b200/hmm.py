"""HMM parameter tables of the reference, built on the host exactly as the R code does (a few dozen doubles; the
Viterbi itself runs in libinfercnv_b200.so).  Used by the Python mirror of the R interface, the benchmark and the tests.

  get_HMM         .get_HMM, R/inferCNV_HMM.R:230-265
  i3HMM_get_HMM   .i3HMM_get_HMM, R/inferCNV_i3HMM.R:99-156
  i3_mean_delta   determine_mean_delta_via_Z, R/inferCNV_i3HMM.R:435-445
"""
from __future__ import annotations

from statistics import NormalDist

import numpy as np

CNV_LEVELS = ["cnv:0.01", "cnv:0.5", "cnv:1", "cnv:1.5", "cnv:2", "cnv:3"]  # R/inferCNV_HMM.R:244-256


def get_HMM(cnv_mean_sd: dict, t: float):
    """(state_transitions, delta, mean[6], sd[6]): diagonal 1 - 5t, off-diagonal t, start in the neutral state."""
    Pi = np.full((6, 6), t, dtype=np.float64, order="F")
    np.fill_diagonal(Pi, 1 - 5 * t)
    delta = np.array([t, t, 1 - 5 * t, t, t, t])
    mean = np.array([cnv_mean_sd[k]["mean"] for k in CNV_LEVELS], dtype=np.float64)
    sd = np.array([cnv_mean_sd[k]["sd"] for k in CNV_LEVELS], dtype=np.float64)
    return Pi, delta, mean, sd


def i3_mean_delta(sigma: float, p: float = 0.05) -> float:
    """|qnorm(p, 0, sigma)|."""
    return abs(NormalDist(0.0, sigma).inv_cdf(p))


def i3HMM_get_HMM(sd_trend: dict, t: float, i3_p_val: float = 0.05, use_KS: bool = False):
    """Three states around mu (diagonal 1 - 5t as written in the reference, not 1 - 2t)."""
    Pi = np.full((3, 3), t, dtype=np.float64, order="F")
    np.fill_diagonal(Pi, 1 - 5 * t)
    delta = np.array([t, 1 - 5 * t, t])
    mu, sigma = sd_trend["mu"], sd_trend["sigma"]
    d = sd_trend["KS_delta"] if use_KS else sd_trend["mean_delta"]
    return Pi, delta, np.array([mu - d, mu, mu + d]), np.array([sigma] * 3)
