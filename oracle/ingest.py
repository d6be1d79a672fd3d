"""CPU oracle for the steps in front of the path (TEST INFRASTRUCTURE ONLY, like the rest of oracle/): the two gene
filters of run() step 2 and counts ingest from a sparse matrix.

* ``.below_min_mean_expr_cutoff``  R/inferCNV_ops.R:2149-2158  ``which(rowMeans(expr) < cutoff)``
* ``require_above_min_cells_ref``  R/inferCNV_ops.R:2177-2209  ``sum(x > 0 & !is.na(x)) >= min_cells_per_gene``
* ``remove_genes``                 R/inferCNV.R:445-457
* ``.normalize_data_matrix_by_seq_depth`` ops.R:3082-3111 (oracle.normalize_by_seq_depth) applied to a
  compressed-sparse-column matrix after densifying it - what R does with a dgCMatrix.

R's ``rowMeans`` accumulates in long double, cell after cell, and rounds the long-double quotient to double; NumPy's
``longdouble`` is the same x87 80-bit type on x86-64.  Pinned by the reference's own known answers
(tests/testthat/test_infer_cnv.R:175-219, tests/test_oracle_ingest.py).
"""
from __future__ import annotations

import numpy as np

from . import oracle as orc


def row_means(X) -> np.ndarray:
    X = np.asarray(X, dtype=np.float64)
    acc = np.zeros(X.shape[0], dtype=np.longdouble)
    for c in range(X.shape[1]):                 # sequential over cells, like do_colsum's row loop
        acc += X[:, c]
    return (acc / np.longdouble(X.shape[1])).astype(np.float64)


def below_min_mean_expr_cutoff(expr_data, min_mean_expr) -> np.ndarray:
    """0-based indices (R returns them 1-based)."""
    return np.flatnonzero(row_means(expr_data) < min_mean_expr)


def n_cells_expressing(expr_data) -> np.ndarray:
    X = np.asarray(expr_data, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return (X > 0).sum(axis=1)


def genes_passing_min_cells(expr_data, min_cells_per_gene) -> np.ndarray:
    return np.flatnonzero(n_cells_expressing(expr_data) >= min_cells_per_gene)


def remove_genes(X, gene_indices_to_remove) -> np.ndarray:
    mask = np.ones(np.asarray(X).shape[0], dtype=bool)
    mask[np.asarray(gene_indices_to_remove, dtype=np.int64)] = False
    return np.asfortranarray(np.asarray(X)[mask])


def csc_to_dense(p, i, x, n_genes) -> np.ndarray:
    p = np.asarray(p)
    D = np.zeros((int(n_genes), len(p) - 1), dtype=np.float64, order="F")
    for c in range(len(p) - 1):
        D[np.asarray(i)[p[c]:p[c + 1]], c] = np.asarray(x)[p[c]:p[c + 1]]
    return D


def ingest_sparse_counts(p, i, x, n_genes, min_mean_expr_cutoff=None, min_cells_per_gene=None, normalize_factor=None):
    """run() steps 2-3 (ops.R:560-586) on the densified matrix, filter after filter as R applies them."""
    D = csc_to_dense(p, i, x, n_genes)
    kept = np.arange(D.shape[0])
    if min_mean_expr_cutoff is not None:
        rm = below_min_mean_expr_cutoff(D, min_mean_expr_cutoff)
        D, kept = remove_genes(D, rm), np.delete(kept, rm)
    if min_cells_per_gene is not None:
        ok = genes_passing_min_cells(D, min_cells_per_gene)
        D, kept = np.asfortranarray(D[ok]), kept[ok]
    with np.errstate(invalid="ignore", divide="ignore"):
        return orc.normalize_by_seq_depth(D, normalize_factor), kept


def scale_infercnv_expr(expr_data) -> np.ndarray:
    """scale_infercnv_expr, R/inferCNV_ops.R:3174-3186: t(scale(t(x))) - scale() centres every column of t(x) (= every
    gene) at its long-double mean and divides by sqrt(sum(centred^2) / (n - 1))."""
    X = np.asarray(expr_data, dtype=np.float64)
    centred = X - row_means(X)[:, None]
    acc = np.zeros(X.shape[0], dtype=np.longdouble)
    for c in range(X.shape[1]):
        acc += centred[:, c].astype(np.longdouble) ** 2
    sd = np.sqrt((acc / max(1, X.shape[1] - 1)).astype(np.float64))
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.asfortranarray(centred / sd[:, None])


def remove_tails(chr_idx, tail_length):
    """.remove_tails, R/inferCNV_ops.R:2370-2385 (indices as given, so 1-based inputs give the reference's answers)."""
    chr_idx = list(chr_idx)
    n = len(chr_idx)
    if tail_length < 3 or n < 3:
        return []
    if n < tail_length * 2:
        tail_length = n // 3
    tail_length = int(tail_length)
    return chr_idx[:tail_length] + chr_idx[n - tail_length:]


def genes_removed_at_ends_of_chromosomes(chr_start, chr_len, window_length):
    """remove_genes_at_ends_of_chromosomes, R/inferCNV_ops.R:3000-3017: 0-based indices to drop."""
    tail = (window_length - 1) / 2
    out = []
    for s, n in zip(chr_start, chr_len):
        out += remove_tails(range(int(s), int(s) + int(n)), tail)
    return np.array(out, dtype=np.int64)
