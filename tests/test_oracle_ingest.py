"""Gene filter oracle (oracle/ingest.py) against the reference's own known answers
(tests/testthat/test_infer_cnv.R:175-219) and its structural properties.  CPU only."""
import numpy as np
import pytest

from oracle import ingest as ori

matrix_one = np.arange(1, 6, dtype=float).reshape(5, 1)                 # matrix(1:5, ncol=1)
matrix_three = np.arange(1, 16, dtype=float).reshape(3, 5).T            # matrix(1:15, ncol=3)


@pytest.mark.parametrize("mat, cutoff, answer", [
    (matrix_one, 10, [1, 2, 3, 4, 5]),        # below_answer_1
    (matrix_three, 10, [1, 2, 3, 4]),         # below_answer_2
    (matrix_one, 2, [1]),                     # below_answer_3
    (matrix_three, 8.4, [1, 2, 3]),           # below_answer_4
    (matrix_one, 0, []),                      # below_answer_5
    (matrix_three, 100, [1, 2, 3, 4, 5]),     # below_answer_6
])
def test_below_min_mean_expr_cutoff_known_answers(mat, cutoff, answer):
    assert (ori.below_min_mean_expr_cutoff(mat, cutoff) + 1).tolist() == answer


def test_row_means_is_the_long_double_quotient():
    rng = np.random.default_rng(0)
    X = rng.poisson(3.0, size=(200, 37)).astype(float)
    m = ori.row_means(X)
    assert np.array_equal(m, (X.sum(axis=1).astype(np.longdouble) / 37).astype(float))   # integer sums are exact
    assert np.allclose(m, X.mean(axis=1), rtol=1e-15, atol=0)


def test_min_cells_filter_and_remove_genes():
    X = np.array([[0, 0, 1], [2, 0, 3], [0, 0, 0], [np.nan, 1, 1], [-1, 4, 0]], dtype=float)
    assert ori.n_cells_expressing(X).tolist() == [1, 2, 0, 2, 1]          # NaN and negatives are not "expressed"
    assert ori.genes_passing_min_cells(X, 2).tolist() == [1, 3]
    assert np.array_equal(ori.remove_genes(np.arange(10.).reshape(5, 2), [0, 3]), [[2, 3], [4, 5], [8, 9]])


def test_sparse_ingest_equals_dense_steps():
    rng = np.random.default_rng(1)
    G, C = 60, 25
    D = rng.poisson(0.4, size=(G, C)).astype(float)
    D[:, 7] = 0                                                            # a cell with no counts at all
    p = np.concatenate([[0], np.cumsum((D != 0).sum(axis=0))])
    i = np.concatenate([np.flatnonzero(D[:, c]) for c in range(C)])
    x = np.concatenate([D[np.flatnonzero(D[:, c]), c] for c in range(C)])
    assert np.array_equal(ori.csc_to_dense(p, i, x, G), D)
    Y, kept = ori.ingest_sparse_counts(p, i, x, G, min_mean_expr_cutoff=0.3, min_cells_per_gene=3, normalize_factor=1e5)
    assert 0 < len(kept) < G and Y.shape == (len(kept), C)
    means = D.mean(axis=1)
    assert set(kept) <= set(np.flatnonzero(means >= 0.3))
    assert np.all(np.isnan(Y[:, 7]))                                       # 0 / 0 in R
    cols = [c for c in range(C) if c != 7]
    assert np.allclose(Y[:, cols].sum(axis=0), 1e5)


def test_remove_tails_known_answers():
    """tests/testthat/test_infer_cnv.R:265-305 (tail_answer_1 .. 5)."""
    assert ori.remove_tails(range(1, 6), 0) == []
    assert ori.remove_tails(range(1, 21), 5) == list(range(1, 6)) + list(range(16, 21))
    assert ori.remove_tails(range(2, 18), 5) == list(range(2, 7)) + list(range(13, 18))
    assert ori.remove_tails(range(5, 16), 5) == list(range(5, 10)) + list(range(11, 16))
    assert ori.remove_tails(range(1, 6), 100) == [1, 5]
    assert ori.genes_removed_at_ends_of_chromosomes([0, 20, 22], [20, 2, 7], 11).tolist() == \
        list(range(0, 5)) + list(range(15, 20)) + [22, 23, 27, 28]


def test_scale_is_the_per_gene_zscore():
    rng = np.random.default_rng(2)
    X = rng.lognormal(0, 1, size=(50, 17))
    Z = ori.scale_infercnv_expr(X)
    assert np.allclose(Z.mean(axis=1), 0, atol=1e-14) and np.allclose(Z.std(axis=1, ddof=1), 1, rtol=1e-13)
    assert np.allclose(Z, (X - X.mean(axis=1, keepdims=True)) / X.std(axis=1, ddof=1, keepdims=True), rtol=1e-12, atol=1e-14)
