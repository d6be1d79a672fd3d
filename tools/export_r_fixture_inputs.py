#!/usr/bin/env python
"""Writes the inputs tools/make_hmm_fixture.R reads (plain binary / text, no R packages needed to parse them): the c1 cells
of tests/golden/hmm_mpmath_c1.npz (smooth-block output of the bundled oligodendroglioma example), the i6 / i3 parameters and
index lists for the median filter.     python tools/export_r_fixture_inputs.py <dir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(out):
    os.makedirs(out, exist_ok=True)
    d = np.load(os.path.join(ROOT, "tests", "golden", "hmm_mpmath_c1.npz"))
    X = np.asfortranarray(d["X"])
    G, C = X.shape
    X.T.astype("<f8").tofile(os.path.join(out, "x.bin"))            # column-major G x C
    np.savetxt(os.path.join(out, "chr_len.txt"), d["chr_len"], fmt="%d")
    meta = [f"G={G}", f"C={C}"]
    for tag in ("i6", "i3"):
        meta.append(f"{tag}_m={len(d[tag + '_mean'])}")
        np.savetxt(os.path.join(out, f"{tag}_Pi.txt"), np.asarray(d[tag + "_Pi"]).T.ravel(), fmt="%.17g")   # column-major for matrix()
        for k in ("delta", "mean", "sd"):
            np.savetxt(os.path.join(out, f"{tag}_{k}.txt"), d[f"{tag}_{k}"], fmt="%.17g")
    open(os.path.join(out, "meta.txt"), "w").write("\n".join(meta) + "\n")
    lists = [list(range(1, 5)), [5, 7, 6, 9, 8, 10]]                 # 1-based: a group of 4 and a permuted group of 6
    open(os.path.join(out, "mf_lists.txt"), "w").write("\n".join(" ".join(map(str, l)) for l in lists) + "\n")
    print(out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r_fixture")
