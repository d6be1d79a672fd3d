"""CNV region calling oracle (oracle/regions.py) against the reference's own bundled known answer and against its
literal transcription.  CPU only."""
import os

import numpy as np
import pytest

from oracle import regions as orr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cnv_regions_fixture.npz")


@pytest.fixture(scope="module")
def fx():
    z = np.load(GOLD)
    codes = z["chr_codes"] - 1                                   # 0-based chromosome codes, rows sorted by chr
    starts = np.flatnonzero(np.concatenate([[True], codes[1:] != codes[:-1]]))
    lens = np.diff(np.concatenate([starts, [len(codes)]]))
    return dict(z=z, codes=codes, chr_start=starts, chr_len=lens, chr_names=[str(s) for s in z["chr_levels"]],
                sub_cells=z["subcluster_cells"] - 1)


def test_literal_transcription_reproduces_the_reference_run(fx):
    """data/HMM_states.rda -> regions == what data/mcmc_obj.rda recorded (generate_cnv_region_reports by
    "subcluster", neutral state 3 ignored): names carry the running region counter, so every region before a
    listed one is checked implicitly."""
    z = fx["z"]
    cons = orr.literal_state_consensus(z["hmm_states"][:, fx["sub_cells"]])
    regions, counter = orr.literal_define_cnv_gene_regions(cons, fx["codes"], fx["chr_names"], 0)
    non_neutral = [(n, g[0] + 1, g[-1] + 1, len(g)) for n, st, g in regions if st != 3]
    want = list(zip(z["region_names"].tolist(), z["region_first_gene"].tolist(), z["region_last_gene"].tolist(),
                    z["region_n_genes"].tolist()))
    assert non_neutral == want
    assert counter == len(regions) >= 22
    # every region's cells are the subcluster's cells (mcmc_obj@cell_gene[[i]]$Cells, stored sorted)
    for row in z["region_cells"]:
        assert sorted((fx["sub_cells"] + 1).tolist()) == row.tolist()


def test_vectorised_restatement_reproduces_the_reference_run(fx):
    z = fx["z"]
    subclusters = {"tumor": {"tumor_s1": fx["sub_cells"]}}
    out = orr.predicted_cnv_regions(z["hmm_states"], fx["chr_start"], fx["chr_len"], fx["chr_names"], z["gene_names"],
                                    z["gene_start"], z["gene_stop"], z["cell_names"], {"normal": z["ref_idx"] - 1},
                                    {"tumor": z["obs_idx"] - 1}, subclusters, by="subcluster")
    assert [g["cell_group_name"] for g in out] == ["tumor.tumor_s1"]
    got = [(n, a + 1, b + 1) for n, st, _, _, _, a, b in out[0]["regions"] if st != 3]
    assert got == list(zip(z["region_names"].tolist(), z["region_first_gene"].tolist(), z["region_last_gene"].tolist()))
    # bounds: min(start) / max(stop) over the region's genes (HMM.R:1078-1079)
    for n, st, ch, lo, hi, a, b in out[0]["regions"]:
        assert lo == z["gene_start"][a:b + 1].min() and hi == z["gene_stop"][a:b + 1].max()
        assert ch == n.split("-region_")[0]


@pytest.mark.parametrize("seed", range(6))
def test_vectorised_equals_literal_on_random_cases(seed):
    rng = np.random.default_rng(seed)
    lens = rng.permutation(np.array([1, 2, 3, 40, 17, 0, 1, 25][: 4 + seed % 5]))
    lens = lens[lens > 0]
    G, C = int(lens.sum()), 12
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    codes = np.repeat(np.arange(len(lens)), lens)
    # few distinct values with long runs -> many ties and many region boundaries
    base = np.repeat(rng.integers(1, 7, size=(G // 5 + 1, C)), 5, axis=0)[:G]
    noise = rng.integers(1, 7, size=(G, C))
    S = np.where(rng.random((G, C)) < 0.3, noise, base).astype(np.uint8)
    if seed % 2:
        S[rng.random((G, C)) < 0.1] = orr.UNASSIGNED_U8               # cells the HMM did not cover (-1 in R)
    gs = np.cumsum(rng.integers(1, 1000, size=G))
    ge = gs + rng.integers(1, 5000, size=G)                             # overlapping genes: max(stop) is not the last one
    names = ["chr%d" % (k + 1) for k in range(len(lens))]
    for cells in (np.arange(C), np.array([3]), np.array([5, 1, 7, 2])):
        cons_l = orr.literal_state_consensus(S[:, cells])
        cons_v = orr.state_consensus(S, cells)
        assert np.array_equal(np.where(cons_l < 0, 255, cons_l), cons_v)
        reg_l, cnt = orr.literal_define_cnv_gene_regions(cons_l, codes, names, 7)
        bounds_l = orr.literal_cnv_gene_region_bounds(reg_l, gs, ge)
        r = orr.cnv_regions(cons_v, starts, lens, gs, ge)
        assert len(r["seq"]) == len(reg_l) == cnt - 7
        for k, ((name, st, genes), (_, _, lo, hi)) in enumerate(zip(reg_l, bounds_l)):
            assert (r["first_gene"][k], r["last_gene"][k]) == (genes[0], genes[-1])
            assert r["state"][k] == st and r["start"][k] == lo and r["end"][k] == hi
            assert name == "%s-region_%d" % (names[r["chr"][k]], 7 + k + 1)


def test_ties_go_to_the_smallest_state():
    m = np.array([[2, 4, 4, 2], [6, 5, 6, 5], [3, 3, 3, 1], [255, 255, 1, 1]], dtype=np.uint8)
    assert orr.literal_state_consensus(m).tolist() == [2, 5, 3, -1]
    assert orr.state_consensus(m, np.arange(4)).tolist() == [2, 5, 3, 255]


def test_report_text_layout():
    S = np.array([[3, 3], [4, 4], [4, 4], [3, 3], [3, 3]], dtype=np.uint8)
    out = orr.predicted_cnv_regions(S, [0, 3], [3, 2], ["chr1", "chr2"], ["A", "B", "C", "D", "E"], [10, 20, 30, 5, 50],
                                    [15, 45, 35, 9, 60], ["c1", "c2"], {"normal": np.array([0])}, {"tumor": np.array([1])},
                                    None, by="cell")                       # no subclusters -> falls back to consensus
    rep = orr.cnv_region_reports(out, ["chr1"] * 3 + ["chr2"] * 2, ["A", "B", "C", "D", "E"], [10, 20, 30, 5, 50],
                                 [15, 45, 35, 9, 60], ignore_neutral_state=3)
    assert rep["cell_groupings"] == "cell_group_name\tcell\nnormal\tc1\ntumor\tc2\n"
    assert rep["pred_cnv_regions.dat"] == ("cell_group_name\tcnv_name\tstate\tchr\tstart\tend\n"
                                           "normal\tchr1-region_2\t4\tchr1\t20\t45\n"
                                           "tumor\tchr1-region_5\t4\tchr1\t20\t45\n")
    assert rep["pred_cnv_genes.dat"].splitlines()[1:] == ["normal\tchr1-region_2\t4\tB\tchr1\t20\t45",
                                                          "normal\tchr1-region_2\t4\tC\tchr1\t30\t35",
                                                          "tumor\tchr1-region_5\t4\tB\tchr1\t20\t45",
                                                          "tumor\tchr1-region_5\t4\tC\tchr1\t30\t35"]
    assert rep["genes_used.dat"].splitlines()[:2] == ["chr\tstart\tstop", "A\tchr1\t10\t15"]
