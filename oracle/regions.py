"""CPU oracle for CNV region calling (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Restates, in NumPy / plain Python, the reference's post-HMM reporting path:

* ``.get_state_consensus``          R/inferCNV_HMM.R:977-988   modal state per gene over a group's cells
* ``.define_cnv_gene_regions``      R/inferCNV_HMM.R:1006-1058 run-length segmentation per chromosome
* ``.get_cnv_gene_region_bounds``   R/inferCNV_HMM.R:1071-1087 (state, chr, min start, max stop) per region
* ``get_predicted_CNV_regions``     R/inferCNV_HMM.R:706-764   groups by "consensus" / "subcluster" / "cell"
* ``generate_cnv_region_reports``   R/inferCNV_HMM.R:790-869   the four tab-separated report files

Pinned by the reference's own bundled data: ``data/HMM_states.rda`` taken through this code reproduces the
region names, gene ranges and cell lists stored in ``data/mcmc_obj.rda`` (tests/golden/cnv_regions_fixture.npz,
tests/test_oracle_regions.py).

Two layers, as in oracle.py: ``literal_*`` = line-by-line transcription of the R loops (small cases),
the rest = vectorised NumPy restatement (what the GPU tests compare against at size).  0-based indices.
"""
from __future__ import annotations

import numpy as np

UNASSIGNED_U8 = 255    # the one-byte wire format's "-1" (include/infercnv_b200.h)


def _as_signed(states) -> np.ndarray:
    """uint8 wire format -> the values R sees (255 -> -1)."""
    s = np.asarray(states)
    if s.dtype == np.uint8:
        s = s.astype(np.int16)
        s[s == UNASSIGNED_U8] = -1
    return s


# ---- literal transcriptions ---------------------------------------------------------------------------

def literal_state_consensus(cell_group_matrix) -> np.ndarray:
    """HMM.R:977-988.  `table(x)` lists the distinct values in increasing order, `order(t, decreasing=TRUE)[1]` is
    the first of the largest counts (radix order is stable), i.e. ties go to the smallest state."""
    m = _as_signed(cell_group_matrix)
    out = np.empty(m.shape[0], dtype=np.int64)
    for i in range(m.shape[0]):
        vals, counts = np.unique(m[i], return_counts=True)      # sorted values, like table()
        best = 0
        for j in range(1, len(vals)):
            if counts[j] > counts[best]:
                best = j
        out[i] = vals[best]
    return out


def literal_define_cnv_gene_regions(state_consensus, chr_of_gene, chr_names, cnv_region_counter):
    """HMM.R:1006-1058.  Returns (ordered list of (region_name, state, [gene indices])), new counter)."""
    regions = []
    chrs = list(dict.fromkeys(np.asarray(chr_of_gene).tolist()))      # unique(gene_order$chr): order of appearance
    for c in chrs:
        gene_idx = np.flatnonzero(np.asarray(chr_of_gene) == c)
        if len(gene_idx) < 2:
            continue
        prev_state = state_consensus[gene_idx[0]]
        cnv_region_counter += 1
        name = "%s-region_%d" % (chr_names[c], cnv_region_counter)
        cur = [int(gene_idx[0])]
        cur_state = prev_state
        for i in gene_idx[1:]:
            state = state_consensus[i]
            if state != prev_state:
                regions.append((name, int(cur_state), cur))
                cnv_region_counter += 1
                name = "%s-region_%d" % (chr_names[c], cnv_region_counter)
                cur, cur_state = [int(i)], state
            else:
                cur.append(int(i))
            prev_state = state
        regions.append((name, int(cur_state), cur))
    return regions, cnv_region_counter


def literal_cnv_gene_region_bounds(regions, gene_start, gene_stop):
    """HMM.R:1071-1087: per region (name, state, min(start), max(stop))."""
    return [(name, state, int(min(gene_start[g] for g in genes)), int(max(gene_stop[g] for g in genes)))
            for name, state, genes in regions]


# ---- vectorised restatement -----------------------------------------------------------------------------

def state_consensus(states, cells) -> np.ndarray:
    """Modal state per gene over the listed cells (ties -> smallest), as uint8 in the wire format."""
    m = _as_signed(np.asarray(states)[:, np.asarray(cells, dtype=np.int64)])
    vals = np.arange(-1, 8)
    counts = np.stack([(m == v).sum(axis=1) for v in vals], axis=1)
    if counts.sum(axis=1).min() != m.shape[1]:
        raise ValueError("state outside -1..7")
    best = vals[np.argmax(counts, axis=1)]                          # argmax returns the first maximum
    return np.where(best < 0, UNASSIGNED_U8, best).astype(np.uint8)


def cnv_regions(seqs, chr_start, chr_len, gene_start, gene_stop):
    """Run-length regions of every column of `seqs` (G x n_seq, uint8), chromosome by chromosome, skipping
    chromosomes with fewer than two genes (HMM.R:1012-1014).  Returns a dict of arrays ordered by
    (sequence, chromosome, position): seq, chr, first_gene, last_gene (inclusive), state, start, end."""
    seqs = np.asarray(seqs)
    if seqs.ndim == 1:
        seqs = seqs[:, None]
    G, n_seq = seqs.shape
    chr_start, chr_len = np.asarray(chr_start, dtype=np.int64), np.asarray(chr_len, dtype=np.int64)
    gene_start, gene_stop = np.asarray(gene_start, dtype=np.float64), np.asarray(gene_stop, dtype=np.float64)
    chr_of = np.repeat(np.arange(len(chr_start)), chr_len)
    valid = np.repeat(chr_len >= 2, chr_len)
    is_chr_first = np.zeros(G, dtype=bool)
    is_chr_first[chr_start[chr_len > 0]] = True
    out = {k: [] for k in ("seq", "chr", "first_gene", "last_gene", "state", "start", "end")}
    for s in range(n_seq):
        col = seqs[:, s]
        flag = valid & (is_chr_first | np.concatenate([[True], col[1:] != col[:-1]]))
        first = np.flatnonzero(flag)
        if len(first) == 0:
            continue
        nxt = np.concatenate([first[1:], [G]])
        chr_end = chr_start[chr_of[first]] + chr_len[chr_of[first]]
        last = np.minimum(nxt, chr_end) - 1
        out["seq"].append(np.full(len(first), s, dtype=np.int32))
        out["chr"].append(chr_of[first].astype(np.int32))
        out["first_gene"].append(first.astype(np.int32))
        out["last_gene"].append(last.astype(np.int32))
        out["state"].append(_as_signed(col[first]).astype(np.int32))
        out["start"].append(np.array([gene_start[a:b + 1].min() for a, b in zip(first, last)]))
        out["end"].append(np.array([gene_stop[a:b + 1].max() for a, b in zip(first, last)]))
    dt = dict(seq=np.int32, chr=np.int32, first_gene=np.int32, last_gene=np.int32, state=np.int32, start=np.float64,
              end=np.float64)
    return {k: (np.concatenate(v) if v else np.zeros(0, dtype=dt[k])).astype(dt[k]) for k, v in out.items()}


def cell_groups_for(by, ref_groups, obs_groups, subclusters, cell_names):
    """HMM.R:709-733: ordered (name, 0-based cell indices) pairs.  `ref_groups` / `obs_groups`: ordered dicts
    name -> indices; `subclusters`: {group: {subcluster: indices}} or None."""
    if subclusters is None:
        by = "consensus"
    if by == "consensus":
        return by, [(n, np.asarray(v)) for n, v in list(ref_groups.items()) + list(obs_groups.items())]
    if by == "subcluster":
        return by, [("%s.%s" % (g, n), np.asarray(v)) for g, sub in subclusters.items() for n, v in sub.items()]
    if by == "cell":
        cells = [int(i) for v in list(ref_groups.values()) + list(obs_groups.values()) for i in v]
        return by, [(cell_names[i], np.array([i])) for i in cells]
    raise ValueError("Error, shouldn't get here ... bug")


def predicted_cnv_regions(states, chr_start, chr_len, chr_names, gene_names, gene_start, gene_stop, cell_names,
                          ref_groups, obs_groups, subclusters, by="consensus"):
    """get_predicted_CNV_regions, HMM.R:706-764.  One entry per cell group:
    dict(cell_group_name, cells (names), regions = [(cnv_name, state, chr_name, start, end, first_gene, last_gene)])."""
    by, groups = cell_groups_for(by, ref_groups, obs_groups, subclusters, cell_names)
    out = []
    counter = 0
    for name, cells in groups:
        cons = state_consensus(states, cells)
        r = cnv_regions(cons, chr_start, chr_len, gene_start, gene_stop)
        regs = []
        for k in range(len(r["seq"])):
            counter += 1
            c = int(r["chr"][k])
            regs.append(("%s-region_%d" % (chr_names[c], counter), int(r["state"][k]), chr_names[c], int(r["start"][k]),
                         int(r["end"][k]), int(r["first_gene"][k]), int(r["last_gene"][k])))
        out.append({"cell_group_name": name, "cells": [cell_names[i] for i in cells], "regions": regs})
    return out


def _fmt_state(v) -> str:
    """write.table on a double column: integers print without a decimal point."""
    return str(int(v)) if float(v) == int(v) else repr(float(v))


def cnv_region_reports(cnv_regions_list, chr_of_gene_names, gene_names, gene_start, gene_stop, ignore_neutral_state=None):
    """generate_cnv_region_reports, HMM.R:790-869: the text of the four files
    (.cell_groupings, .pred_cnv_regions.dat, .pred_cnv_genes.dat, .genes_used.dat), as write.table(quote=FALSE,
    sep="\\t") prints them (row.names=FALSE for the first three; the gene order file keeps its row names, so its
    header has one field fewer than its rows)."""
    keep = (lambda st: True) if ignore_neutral_state is None else (lambda st: st != ignore_neutral_state)
    lines = ["cell_group_name\tcell"]
    for g in cnv_regions_list:
        lines += ["%s\t%s" % (g["cell_group_name"], c) for c in g["cells"]]
    groupings = "\n".join(lines) + "\n"
    lines = ["cell_group_name\tcnv_name\tstate\tchr\tstart\tend"]
    for g in cnv_regions_list:
        lines += ["%s\t%s\t%s\t%s\t%d\t%d" % (g["cell_group_name"], n, _fmt_state(st), ch, a, b)
                  for n, st, ch, a, b, _, _ in g["regions"] if keep(st)]
    regions = "\n".join(lines) + "\n"
    lines = ["cell_group_name\tgene_region_name\tstate\tgene\tchr\tstart\tend"]
    for g in cnv_regions_list:
        for n, st, ch, _, _, first, last in g["regions"]:
            if keep(st):
                lines += ["%s\t%s\t%s\t%s\t%s\t%d\t%d" % (g["cell_group_name"], n, _fmt_state(st), gene_names[i], ch,
                                                         gene_start[i], gene_stop[i]) for i in range(first, last + 1)]
    genes = "\n".join(lines) + "\n"
    lines = ["chr\tstart\tstop"] + ["%s\t%s\t%d\t%d" % (gene_names[i], chr_of_gene_names[i], gene_start[i], gene_stop[i])
                                    for i in range(len(gene_names))]
    genes_used = "\n".join(lines) + "\n"
    return {"cell_groupings": groupings, "pred_cnv_regions.dat": regions, "pred_cnv_genes.dat": genes,
            "genes_used.dat": genes_used}
