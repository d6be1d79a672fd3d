#!/usr/bin/env python
"""Independent pin of the HMM state calls: Viterbi.dthmm.adj (R/inferCNV_HMM.R:1101-1176) restated in 50-digit
arithmetic (mpmath), run on real data, with the smallest decision margin of every sequence.

The reference holds no HMM fixture and R is not available in the build image, so the C oracle's Viterbi is otherwise
checked against a Python transcription only.  This script removes the "same floating point" caveat:
it evaluates log Q(z) = log(erfc(z / sqrt 2) / 2), the normalised emissions, the recursion and the trace-back at 50
significant digits and records, per sequence, the smallest gap between the winner and the runner-up of ANY arg-max taken
(forward maxima, final state, trace-back).  A margin far above double-precision rounding (~1e-12 accumulated over a
chromosome) means the double-precision answer cannot differ from the exact one; tests/test_oracle_hmm.py and the GPU
parity tests compare their states with the ones stored here.

Input: the bundled oligodendroglioma example (config c1) through the oracle's smooth block; cells = a few of every
reference / observation group; every chromosome.  i6 with the fixture means / sds (data/mcmc_obj.rda), t = 1e-6, and i3
with mu / sigma from the reference cells.

    python tools/make_hmm_mpmath_fixture.py      # ~2-4 minutes; writes tests/golden/hmm_mpmath_c1.npz
"""
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

mp.mp.dps = 50


def split(idx, off):
    return [idx[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def viterbi_mp(x, Pi, delta, mean, sd):
    """-> (states 1..m, smallest arg-max margin as float).  x: doubles (exact inputs), parameters doubles."""
    n, m = len(x), len(mean)
    if n < 2:
        return [3] * n, float("inf")
    s = sorted(mp.mpf(float(v)) for v in sd)
    sdm = s[m // 2] if m % 2 else (s[m // 2 - 1] + s[m // 2]) / 2          # median(sd), HMM.R:1122
    mean = [mp.mpf(float(v)) for v in mean]
    logPi = [[mp.log(mp.mpf(float(Pi[j, k]))) for k in range(m)] for j in range(m)]
    sqrt2 = mp.sqrt(2)

    def emis(xi):
        e = []
        for k in range(m):
            z = abs(mp.mpf(float(xi)) - mean[k]) / sdm
            logq = mp.log(mp.erfc(z / sqrt2) / 2)                              # pnorm(z, log.p=TRUE, lower.tail=FALSE)
            e.append(1 / (-logq))
        tot = sum(e)
        return [mp.log(v / tot) for v in e]

    margin = mp.inf
    nu = [[None] * m for _ in range(n)]
    le = emis(x[0])
    nu[0] = [mp.log(mp.mpf(float(delta[k]))) + le[k] for k in range(m)]
    for i in range(1, n):
        le = emis(x[i])
        for k in range(m):
            c = sorted((nu[i - 1][j] + logPi[j][k] for j in range(m)), reverse=True)
            margin = min(margin, c[0] - c[1])
            nu[i][k] = c[0] + le[k]
    y = [0] * n
    c = sorted(nu[n - 1], reverse=True)
    margin = min(margin, c[0] - c[1])
    y[n - 1] = max(range(m), key=lambda k: (nu[n - 1][k], -k))
    for i in range(n - 2, -1, -1):
        cand = [logPi[j][y[i + 1]] + nu[i][j] for j in range(m)]
        c = sorted(cand, reverse=True)
        margin = min(margin, c[0] - c[1])
        y[i] = max(range(m), key=lambda j: (cand[j], -j))
    return [v + 1 for v in y], float(margin)


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "oligodendroglioma.npz"))
    h = np.load(os.path.join(ROOT, "tests", "golden", "hmm_fixture.npz"))
    counts = np.asfortranarray(d["counts"].astype(np.float64))
    cs, cl = orc.chr_ranges(d["chr_codes"])
    refs = [g - 1 for g in split(d["ref_idx"], d["ref_off"])]
    obs = [g - 1 for g in split(d["obs_idx"], d["obs_off"])]
    S = orc.smooth_block(orc.normalize_by_seq_depth(counts), cs, cl, refs, nthreads=orc.max_threads())
    cells = np.array([g[0] for g in refs] + [g[i] for g in obs for i in (0, len(g) // 2)], dtype=np.int64)   # 2 + 8 cells
    X = np.asfortranarray(S[:, cells])
    out = {"cells": cells, "X": X, "chr_start": cs, "chr_len": cl}
    Pi6, delta6 = orc.hmm_params(6)
    mu, sg = orc.mean_sd_over_cells(S, np.concatenate(refs))
    Pi3, d3, mean3, sd3 = orc.i3_hmm_params(S, np.concatenate(refs))
    for name, Pi, delta, mean, sd in (("i6", Pi6, delta6, h["mu"], h["sd"]), ("i3", Pi3, d3, mean3, sd3)):
        states = np.zeros(X.shape, dtype=np.uint8)
        margins = np.zeros((len(cs), X.shape[1]))
        for c in range(X.shape[1]):
            for k, (s0, n) in enumerate(zip(cs, cl)):
                y, mg = viterbi_mp(X[s0:s0 + n, c], Pi, delta, mean, sd)
                states[s0:s0 + n, c] = y
                margins[k, c] = mg
            print(f"{name}: cell {c + 1}/{X.shape[1]} done, min margin so far {margins[:, :c + 1].min():.3e}", file=sys.stderr)
        out[f"{name}_states"] = states
        out[f"{name}_margins"] = margins
        out[f"{name}_Pi"], out[f"{name}_delta"], out[f"{name}_mean"], out[f"{name}_sd"] = Pi, delta, np.asarray(mean), np.asarray(sd)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hmm_mpmath_c1.npz"), **out)
    print("i6 min margin %.3e, i3 min margin %.3e" % (out["i6_margins"].min(), out["i3_margins"].min()))


if __name__ == "__main__":
    main()
