#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's own bundled data.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [/root/reference]

Outputs (all under tests/golden/, committed):

* example_object.npz   - `data/infercnv_object_example.rda`: `count.data` (input) and `expr.data`
                         (the reference's own output of run() steps 1-15 + denoise; options
                         cutoff=1, cluster_by_groups=TRUE, denoise=TRUE, HMM=FALSE), chromosome
                         codes, reference/observation cell indices, subclusters.
                         This is the primary known-answer test for the whole smooth block.
* hmm_fixture.npz      - `data/mcmc_obj.rda@mu/@sig` (six i6 emission means / precisions) and
                         `data/HMM_states.rda` (RNG-dependent run, NOT a strict golden; kept for
                         reference statistics only).
* oligodendroglioma.npz- `inst/extdata` example of `example/run.R` taken through the reference's
                         ingest rules (R/inferCNV.R:133-337 gene ordering + chr_exclude + >=100
                         counts/cell; R/inferCNV_ops.R:2128-2213 cutoff=1, min_cells_per_gene=3)
                         -> 8508 genes x 184 cells.  Stored as the filtered raw count matrix so
                         both the oracle and the CUDA path start from identical bytes.

* cnv_regions_fixture.npz - the reference's own known answer for CNV region calling (R/inferCNV_HMM.R:706-1087):
                         `data/HMM_states.rda` (state matrix, 4613 x 20) is the input; `data/mcmc_obj.rda`
                         carries what `generate_cnv_region_reports(by="subcluster")` made of it in the
                         reference's run - `@cnv_regions` (region names, numbered by the running region
                         counter), `@cell_gene[[i]]$Genes` / `$Cells` (the genes and cells of every
                         non-neutral region, read back from `.pred_cnv_genes.dat` /
                         `.cell_groupings`), plus the `gene_order` and `tumor_subclusters` it ran with.

Nothing here is reference source code: it is the reference's *data*, reduced to the matrices the
hot path consumes, plus this script that made them.
"""
from __future__ import annotations

import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from rda_reader import RObj, as_factor, as_matrix, named_list, read_rda  # noqa: E402


def example_object(ref: str) -> None:
    d = read_rda(os.path.join(ref, "data/infercnv_object_example.rda"))
    o = d["infercnv_object_example"]
    expr = as_matrix(o.attr["expr.data"])
    counts = as_matrix(o.attr["count.data"])
    go = named_list(o.attr["gene_order"])
    chr_codes, chr_levels = as_factor(go["chr"])
    refs = named_list(o.attr["reference_grouped_cell_indices"])
    obs = named_list(o.attr["observation_grouped_cell_indices"])
    sub = named_list(named_list(o.attr["tumor_subclusters"])["subclusters"])
    sub_flat = {}
    for grp, lst in sub.items():
        for name, idx in named_list(lst).items():
            sub_flat[f"{grp}/{name}"] = np.asarray(idx.value, dtype=np.int32)
    assert expr.shape == counts.shape == (4613, 20)
    np.savez_compressed(
        os.path.join(HERE, "example_object.npz"),
        counts=np.asfortranarray(counts.astype(np.int32)),
        expr=np.asfortranarray(expr),
        chr_codes=chr_codes.astype(np.int32),
        chr_levels=np.array(chr_levels),
        ref_names=np.array(list(refs)),
        ref_idx=np.concatenate([np.asarray(v.value, dtype=np.int32) for v in refs.values()]),
        ref_off=np.cumsum([0] + [len(v.value) for v in refs.values()]).astype(np.int32),
        obs_names=np.array(list(obs)),
        obs_idx=np.concatenate([np.asarray(v.value, dtype=np.int32) for v in obs.values()]),
        obs_off=np.cumsum([0] + [len(v.value) for v in obs.values()]).astype(np.int32),
        sub_names=np.array(list(sub_flat)),
        sub_idx=np.concatenate(list(sub_flat.values())),
        sub_off=np.cumsum([0] + [len(v) for v in sub_flat.values()]).astype(np.int32),
    )
    print("example_object.npz", expr.shape, "chr levels", len(chr_levels))


def hmm_fixture(ref: str) -> None:
    m = read_rda(os.path.join(ref, "data/mcmc_obj.rda"))["mcmc_obj"]
    mu = np.asarray(m.attr["mu"].value, dtype=np.float64)
    sig = np.asarray(m.attr["sig"].value, dtype=np.float64)
    h = read_rda(os.path.join(ref, "data/HMM_states.rda"))["HMM_states"]
    states = as_matrix(h)
    np.savez_compressed(
        os.path.join(HERE, "hmm_fixture.npz"),
        mu=mu, sig=sig, sd=1.0 / np.sqrt(sig),
        hmm_states=np.asfortranarray(states.astype(np.int8)),
    )
    print("hmm_fixture.npz mu", mu, "sd", 1.0 / np.sqrt(sig), "states", states.shape)


def cnv_regions_fixture(ref: str) -> None:
    m = read_rda(os.path.join(ref, "data/mcmc_obj.rda"))["mcmc_obj"]
    h = read_rda(os.path.join(ref, "data/HMM_states.rda"))["HMM_states"]
    states = as_matrix(h)
    gene_names, cell_names = (np.array(v.value) for v in h.attr["dimnames"].value)
    go = m.attr["gene_order"]
    chr_f = go.value[0]
    assert list(go.attr["row.names"].value) == list(gene_names)
    sub = m.attr["tumor_subclusters"].value[1]          # $subclusters: list(tumor = list(tumor_s1 = <cells>))
    sub_names, sub_cells = [], []
    for gname, grp in zip(sub.attr["names"].value, sub.value):
        for sname, sc in zip(grp.attr["names"].value, grp.value):
            sub_names.append(f"{gname}.{sname}")          # unlist(recursive=FALSE) naming, HMM.R:718
            sub_cells.append(np.asarray(sc.value, dtype=np.int32))
    levels = m.attr["cnv_regions"].attr["levels"].value
    names, first, last, count, cells = [], [], [], [], []
    for cg in m.attr["cell_gene"].value:
        reg, genes, cl = cg.value
        names.append(levels[int(reg.value[0]) - 1])
        g = np.asarray(genes.value)
        assert np.array_equal(g, np.arange(g[0], g[-1] + 1)), "a region is a run of consecutive genes"
        first.append(g[0]); last.append(g[-1]); count.append(len(g))
        cells.append(np.asarray(cl.value, dtype=np.int32))
    np.savez_compressed(
        os.path.join(HERE, "cnv_regions_fixture.npz"),
        hmm_states=np.asfortranarray(states.astype(np.uint8)),
        gene_names=gene_names, cell_names=cell_names,
        chr_codes=np.asarray(chr_f.value, dtype=np.int32), chr_levels=np.array(chr_f.attr["levels"].value),
        gene_start=np.asarray(go.value[1].value, dtype=np.int64), gene_stop=np.asarray(go.value[2].value, dtype=np.int64),
        subcluster_names=np.array(sub_names), subcluster_cells=np.concatenate(sub_cells),      # 1-based, as in R
        subcluster_off=np.cumsum([0] + [len(v) for v in sub_cells]).astype(np.int32),
        ref_idx=np.asarray(m.attr["reference_grouped_cell_indices"].value[0].value, dtype=np.int32),
        obs_idx=np.asarray(m.attr["observation_grouped_cell_indices"].value[0].value, dtype=np.int32),
        region_names=np.array(names), region_first_gene=np.array(first, dtype=np.int32),       # 1-based, as in R
        region_last_gene=np.array(last, dtype=np.int32), region_n_genes=np.array(count, dtype=np.int32),
        region_cells=np.stack(cells),
    )
    print("cnv_regions_fixture.npz states", states.shape, "non-neutral regions", names)


def oligodendroglioma(ref: str) -> None:
    ext = os.path.join(ref, "inst/extdata")
    with gzip.open(os.path.join(ext, "oligodendroglioma_expression_downsampled.counts.matrix.gz"), "rt") as f:
        cells = f.readline().rstrip("\n").split("\t")
        genes, rows = [], []
        for line in f:
            parts = line.rstrip("\n").split("\t")
            genes.append(parts[0])
            rows.append(np.array(parts[1:], dtype=np.float64))
    raw = np.vstack(rows)
    assert raw.shape == (len(genes), len(cells))
    # gene order file (R/inferCNV.R:163-183): name, chr, start, stop; drop chr_exclude
    pos_names, pos_chr, pos_start, pos_stop = [], [], [], []
    for line in open(os.path.join(ext, "gencode_downsampled.EXAMPLE_ONLY_DONT_REUSE.txt")):
        a = line.rstrip("\n").split("\t")
        if a[1] in ("chrX", "chrY", "chrM"):
            continue
        pos_names.append(a[0]); pos_chr.append(a[1]); pos_start.append(int(a[2])); pos_stop.append(int(a[3]))
    # annotations (R/inferCNV.R:186-196)
    ann = {}
    ann_order = []
    for line in open(os.path.join(ext, "oligodendroglioma_annotations_downsampled.txt")):
        a = line.rstrip("\n").split("\t")
        ann[a[0]] = a[1]
        ann_order.append(a[0])
    # .order_reduce (R/inferCNV.R:352-428): intersect, chr factor levels in order of appearance,
    # order(chr, start, stop)
    chr_levels = list(dict.fromkeys(pos_chr))
    lvl = {c: i for i, c in enumerate(chr_levels)}
    pos_index = {}
    for i, n in enumerate(pos_names):
        pos_index.setdefault(n, i)
    gene_row = {}
    for i, g in enumerate(genes):
        gene_row.setdefault(g, i)
    keep = [g for g in dict.fromkeys(genes) if g in pos_index]
    keys = sorted(keep, key=lambda g: (lvl[pos_chr[pos_index[g]]], pos_start[pos_index[g]], pos_stop[pos_index[g]]))
    # R's order() is stable: ties keep `keep` order; python's sorted is stable too.
    mat = raw[[gene_row[g] for g in keys], :]
    chr_of = [pos_chr[pos_index[g]] for g in keys]
    # drop unused levels (droplevels, R/inferCNV.R:236)
    used = [c for c in chr_levels if c in set(chr_of)]
    code = {c: i + 1 for i, c in enumerate(used)}
    chr_codes = np.array([code[c] for c in chr_of], dtype=np.int32)
    # cells: >= 100 counts (R/inferCNV.R:252-262), then restrict to annotated cells
    cs = mat.sum(axis=0)
    keep_cells = [j for j in range(len(cells)) if cs[j] >= 100 and cells[j] in ann]
    mat = mat[:, keep_cells]
    cell_names = [cells[j] for j in keep_cells]
    classes = [ann[c] for c in cell_names]
    # run() step 2 (R/inferCNV_ops.R:2128-2213): cutoff=1 on rowMeans, then >=3 cells with x>0
    keep_g = mat.mean(axis=1) >= 1.0
    mat, chr_codes = mat[keep_g], chr_codes[keep_g]
    keep_g2 = (mat > 0).sum(axis=1) >= 3
    mat, chr_codes = mat[keep_g2], chr_codes[keep_g2]
    ref_names = ["Microglia/Macrophage", "Oligodendrocytes (non-malignant)"]
    obs_names = sorted(set(classes) - set(ref_names))

    def idx(name):
        return np.array([i + 1 for i, c in enumerate(classes) if c == name], dtype=np.int32)

    ref = [idx(n) for n in ref_names]
    obs = [idx(n) for n in obs_names]
    as32 = mat.astype(np.float32)
    store = as32 if np.array_equal(as32.astype(np.float64), mat) else mat
    np.savez_compressed(
        os.path.join(HERE, "oligodendroglioma.npz"),
        counts=np.asfortranarray(store),
        chr_codes=chr_codes,
        chr_levels=np.array(used),
        ref_names=np.array(ref_names), ref_idx=np.concatenate(ref),
        ref_off=np.cumsum([0] + [len(v) for v in ref]).astype(np.int32),
        obs_names=np.array(obs_names), obs_idx=np.concatenate(obs),
        obs_off=np.cumsum([0] + [len(v) for v in obs]).astype(np.int32),
    )
    lens = np.bincount(chr_codes)[1:]
    print("oligodendroglioma.npz", mat.shape, store.dtype, "chr lengths", lens.tolist())
    print("  groups", {n: len(v) for n, v in zip(ref_names + obs_names, ref + obs)})


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    example_object(ref)
    hmm_fixture(ref)
    cnv_regions_fixture(ref)
    oligodendroglioma(ref)
