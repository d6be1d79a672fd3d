#!/usr/bin/env python
"""TEST INFRASTRUCTURE: differential fuzzing of the library against the oracle under the host emulation - random small
shapes (single genes, single cells, one-gene chromosomes, windows longer than a chromosome, constant columns, ties,
groups of one) through the public entry points.  `python tests/host/fuzz_emulated.py [n_cases] [seed]`."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import build_emu  # noqa: E402
import numpy as np  # noqa: E402

from infercnv_b200 import _lib  # noqa: E402

_lib.LIB_PATH = build_emu.build()

from infercnv_b200 import api  # noqa: E402
from oracle import denoise as ord_  # noqa: E402
from oracle import ingest as ori  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import regions as orr  # noqa: E402

I6_MEAN = np.array([0.41234766, 0.84075773, 1.01693983, 1.12238786, 1.23842619, 1.44298781])
I6_SD = np.array([0.028893, 0.164549, 0.105553, 0.190574, 0.244093, 0.290072])


def layout(rng):
    K = int(rng.integers(1, 6))
    kind = rng.integers(0, 4)
    if kind == 0:
        lens = rng.integers(1, 4, size=K)
    elif kind == 1:
        lens = rng.integers(1, 40, size=K)
    elif kind == 2:
        lens = rng.integers(90, 260, size=K)
    else:
        lens = np.where(rng.random(K) < 0.4, 1, rng.integers(2, 130, size=K))
    lens = lens.astype(np.int32)
    return np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32), lens


def matrix(rng, G, C):
    kind = rng.integers(0, 5)
    if kind == 0:
        X = rng.poisson(rng.lognormal(0.5, 1.0, size=(G, 1)) * rng.lognormal(0, 0.2, size=(1, C))).astype(float)
    elif kind == 1:
        X = np.full((G, C), float(rng.integers(0, 4)))                 # constant: zero variance everywhere
    elif kind == 2:
        X = rng.integers(0, 3, size=(G, C)).astype(float)              # few distinct values: ties in every median
    elif kind == 3:
        X = rng.lognormal(0, 2.0, size=(G, C))
    else:
        X = rng.poisson(0.2, size=(G, C)).astype(float)                # mostly zeros
    return np.asfortranarray(X)


def groups_of(rng, C):
    n = int(rng.integers(1, min(C, 3) + 1))
    perm = rng.permutation(C)
    cuts = sorted(rng.choice(np.arange(1, C), size=n - 1, replace=False).tolist()) if C > 1 and n > 1 else []
    return [g for g in np.split(perm[: max(n, int(rng.integers(n, C + 1)))], [c for c in cuts if c < C]) if len(g)]


def one_case(rng, case):
    cs, cl = layout(rng)
    G, C = int(cl.sum()), int(rng.integers(1, 9))
    X = matrix(rng, G, C)
    refs = groups_of(rng, C)[:2]
    window = int(rng.choice([3, 5, 11, 51, 101, 151]))
    what = []
    # smooth block (fused) and its pieces
    got = api.smooth_block(X, cs, cl, refs, apply_log=True, threshold=3.0, window_length=window)
    want = orc.smooth_block(X, cs, cl, refs, window=window)
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert err < 1e-10, ("smooth_block", case, G, C, window, err)
    what.append("block")
    Y = api.smooth(X, cs, cl, window)
    # the O(1) smoother differences second-order prefix sums Q ~ n * sum|x| over a chromosome: |error| ~ ulp(Q) / D, far
    # inside the 1e-5 of the north star but not 1e-12 relative on heavy-tailed stand-alone inputs with short windows
    assert np.allclose(Y, orc.smooth_by_chromosome(X, cs, cl, window), rtol=1e-12, atol=1e-15 * cl.max() * np.abs(X).sum(axis=0).max()), ("smooth", case)
    for method in ("median", "mean"):
        assert np.allclose(api.center(X, method), orc.center_columns(X, method), rtol=1e-13, atol=1e-13), ("center", method, case)
    # HMM on the smoothed matrix, per cell and per group, i6 and i3, both arithmetic modes
    S = np.asfortranarray(np.clip(got, 0.05, 5.0))
    for m in (6, 3):
        Pi, delta = orc.hmm_params(m)
        mean, sd = (I6_MEAN, I6_SD) if m == 6 else (np.array([0.8, 1.0, 1.2]), np.array([0.1, 0.1, 0.1]))
        want_st = orc.viterbi_matrix(S, cs, cl, Pi, delta, mean, sd)
        for mode in (0, 1):
            api.set_hmm_mode(mode)
            assert np.array_equal(api.viterbi(S, cs, cl, Pi, delta, mean, sd), want_st), ("viterbi", m, mode, case, G, C)
        grp = groups_of(rng, C)
        sds = np.tile(sd, len(grp)) * rng.uniform(0.5, 1.5)
        want_g = orc.viterbi_matrix(S, cs, cl, Pi, delta, mean, sds, groups=grp)
        assert np.array_equal(api.viterbi(S, cs, cl, Pi, delta, mean, sds, groups=grp), want_g), ("viterbi groups", m, case)
    api.set_hmm_mode(1)
    # median filter over random index lists
    lists = groups_of(rng, C)
    for ws in (3, 7):
        mf = api.median_filter(S, cs, cl, lists, ws)
        assert np.allclose(mf, orc.median_filter(S, cs, cl, lists, ws), rtol=0, atol=1e-15), ("median_filter", ws, case)
    # region calling on the states
    st8 = np.asfortranarray(want_st.astype(np.uint8))
    gs = np.cumsum(rng.integers(1, 100, size=G)).astype(float)
    ge = gs + rng.integers(1, 500, size=G)
    reg, cons = api.predicted_cnv_regions(st8, cs, cl, gs, ge, lists, want_consensus=True)
    want_cons = np.stack([orr.state_consensus(st8, g) for g in lists], axis=1)
    assert np.array_equal(cons, want_cons), ("consensus", case)
    w = orr.cnv_regions(want_cons, cs, cl, gs, ge)
    for k in w:
        assert np.array_equal(reg[k], w[k]), ("regions", k, case)
    return G, C


def _csc(D):
    p = np.concatenate([[0], np.cumsum((D != 0).sum(axis=0))]).astype(np.int32)
    i = np.concatenate([np.flatnonzero(D[:, c]) for c in range(D.shape[1])] + [np.zeros(0, int)]).astype(np.int32)
    x = np.concatenate([D[np.flatnonzero(D[:, c]), c] for c in range(D.shape[1])] + [np.zeros(0)]).astype(np.float64)
    return p, i, x


def widened_case(rng, case):
    """gene statistics (dense and sparse), row selection, sparse normalisation, outlier clamp, scaling, noise clearing,
    median-filter windows up to 21, on shapes down to one gene / one cell with empty rows and columns."""
    G, C = int(rng.integers(1, 70)), int(rng.integers(1, 90))
    D = rng.poisson(rng.choice([0.05, 0.5, 3.0]), size=(G, C)).astype(float)
    if rng.random() < 0.3:
        D[:, rng.integers(0, C)] = 0
    if rng.random() < 0.3:
        D[rng.integers(0, G)] = 0
    sums, npos, means = api.gene_stats(D)
    assert np.array_equal(sums, D.sum(axis=1)) and np.array_equal(npos, ori.n_cells_expressing(D)), ("gene_stats", case)
    assert np.array_equal(means, ori.row_means(D)), ("row means", case)
    p, i, x = _csc(D)
    s2, n2, m2 = api.csc_gene_stats(p, i, x, G)
    assert np.array_equal(s2, sums) and np.array_equal(n2, npos) and np.array_equal(m2, means), ("csc stats", case)
    keep = np.flatnonzero(rng.random(G) < 0.7)
    keep = keep if len(keep) else np.array([0])
    assert np.array_equal(api.remove_genes(D, keep), D[keep]), ("remove_genes", case)
    perm = rng.integers(0, G, size=int(rng.integers(1, 2 * G + 1)))
    assert np.array_equal(api.gather_genes(D, perm), D[perm]), ("gather_genes", case)
    for nf in (None, 1e4):
        Y, cs_ = api.csc_normalize(p, i, x, G, keep=keep, normalize_factor=nf, want_col_sums=True)
        Dk = np.asfortranarray(D[keep])
        with np.errstate(all="ignore"):
            want = orc.normalize_by_seq_depth(Dk, nf)
        assert np.array_equal(cs_, Dk.sum(axis=0)) and np.array_equal(np.isnan(Y), np.isnan(want)), ("csc_normalize", case)
        assert np.array_equal(Y[~np.isnan(want)], want[~np.isnan(want)]), ("csc_normalize values", case, nf)
    E = np.asfortranarray(rng.lognormal(0, 0.3, size=(max(G, 2), C)))
    got, b = api.remove_outliers_norm(E, want_bounds=True)
    assert b == ord_.get_average_bounds(E) and np.array_equal(got, ord_.remove_outliers_norm(E)), ("outliers", case)
    if C >= 2:
        assert np.allclose(api.scale_infercnv_expr(E), ori.scale_infercnv_expr(E), rtol=1e-10, atol=1e-12), ("scale", case)
    refs = np.flatnonzero(rng.random(C) < 0.5)
    thr = float(rng.choice([0.05, 0.3]))
    g1, w1 = api.clear_noise(E, refs, thr), ord_.clear_noise(E, refs if len(refs) else None, thr)
    assert np.mean(np.abs(g1 - w1) > 1e-13 * np.abs(w1)) < 0.02, ("clear_noise", case)
    cs, cl = layout(rng)
    X = matrix(rng, int(cl.sum()), C)
    lists = groups_of(rng, C)
    for ws in (5, 9, 11, 21):
        assert np.allclose(api.median_filter(X, cs, cl, lists, ws), orc.median_filter(X, cs, cl, lists, ws), rtol=0,
                           atol=1e-15 * max(1.0, np.abs(X).max())), ("median_filter", ws, case)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    api.init(0)
    shapes = []
    for case in range(n):
        shapes.append(one_case(rng, case))
        if case % 2 == 0:
            widened_case(rng, case)
    print(f"{n} cases ok (seed {seed}); genes {min(s[0] for s in shapes)}..{max(s[0] for s in shapes)}, "
          f"cells {min(s[1] for s in shapes)}..{max(s[1] for s in shapes)}")


if __name__ == "__main__":
    main()
