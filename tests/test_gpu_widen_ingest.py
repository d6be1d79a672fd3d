"""run() steps 2-3 (gene filters, gene removal, counts ingest incl. sparse input) through
the C ABI against the oracle and the reference's known answers (tests/testthat/test_infer_cnv.R:175-219)."""
import numpy as np
import pytest

from oracle import ingest as ori
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

matrix_one = np.arange(1, 6, dtype=float).reshape(5, 1)
matrix_three = np.arange(1, 16, dtype=float).reshape(3, 5).T


def _csc(D):
    p = np.concatenate([[0], np.cumsum((D != 0).sum(axis=0))]).astype(np.int32)
    i = np.concatenate([np.flatnonzero(D[:, c]) for c in range(D.shape[1])]).astype(np.int32)
    x = np.concatenate([D[np.flatnonzero(D[:, c]), c] for c in range(D.shape[1])]).astype(np.float64)
    return p, i, x


@pytest.mark.parametrize("mat, cutoff, answer", [
    (matrix_one, 10, [1, 2, 3, 4, 5]), (matrix_three, 10, [1, 2, 3, 4]), (matrix_one, 2, [1]),
    (matrix_three, 8.4, [1, 2, 3]), (matrix_one, 0, []), (matrix_three, 100, [1, 2, 3, 4, 5]),
])
def test_below_min_mean_expr_cutoff_known_answers(mat, cutoff, answer):
    from mirror import ops
    assert (ops.below_min_mean_expr_cutoff(mat, cutoff) + 1).tolist() == answer


@pytest.mark.parametrize("shape", [(4613, 20), (257, 1000), (1000, 33), (3, 70000)])
def test_gene_stats_match_the_oracle(shape):
    from infercnv_b200 import api
    rng = np.random.default_rng(shape[0])
    G, C = shape
    X = rng.poisson(rng.lognormal(-1.0, 1.5, size=(G, 1)) * np.ones((1, C))).astype(np.float64)
    sums, n_pos, means = api.gene_stats(X)
    assert np.array_equal(sums, X.sum(axis=1))                 # count data: every order of summation is exact
    assert np.array_equal(n_pos, ori.n_cells_expressing(X))
    assert np.array_equal(means, ori.row_means(X))             # bit-identical to rowMeans
    # non-integer data: same up to the rounding of a different summation order
    Xn = X * rng.random(X.shape)
    _, n_pos, means = api.gene_stats(Xn)
    assert np.array_equal(n_pos, ori.n_cells_expressing(Xn))
    assert np.allclose(means, ori.row_means(Xn), rtol=1e-13, atol=1e-300)


def test_filters_and_remove_genes_on_the_reference_example(oligo):
    """the reference's bundled oligodendroglioma counts, thresholds of example/run.R (cutoff=1) and run()'s default
    min_cells_per_gene=3: same genes kept as the oracle, rows removed from every slot."""
    from mirror import ops
    X = oligo["counts"]
    G, C = X.shape
    rng = np.random.default_rng(0)
    X = np.asfortranarray(np.where(rng.random(X.shape) < 0.35, 0.0, X))        # thin it out so both filters bite
    obj = ops.Infercnv(expr_data=X, count_data=X.copy(order="F"), gene_order_chr=oligo["chr_codes"],
                       gene_names=["g%d" % i for i in range(G)], gene_order_start=np.arange(G), gene_order_stop=np.arange(G) + 9)
    out = ops.require_above_min_mean_expr_cutoff(obj, 1.0)
    rm = ori.below_min_mean_expr_cutoff(X, 1.0)
    assert 0 < len(rm) < G
    want = ori.remove_genes(X, rm)
    assert np.array_equal(out.expr_data, want) and np.array_equal(out.count_data, want)
    kept = np.delete(np.arange(G), rm)
    assert out.gene_names == ["g%d" % i for i in kept] and np.array_equal(out.gene_order_chr, oligo["chr_codes"][kept])
    out2 = ops.require_above_min_cells_ref(out, 60)
    ok = ori.genes_passing_min_cells(want, 60)
    assert 0 < len(ok) < len(kept)
    assert np.array_equal(out2.expr_data, want[ok]) and np.array_equal(out2.gene_order_start, kept[ok])
    with pytest.raises(RuntimeError):
        ops.require_above_min_cells_ref(out, C + 1)                             # stop(998): all genes removed


def test_sparse_counts_ingest_matches_the_dense_reference_steps(oligo):
    from infercnv_b200 import api
    from mirror import ops
    D = np.rint(oligo["counts"])                 # whole counts: sums are exact in any order -> bit-identical checks
    rng = np.random.default_rng(2)
    D[rng.random(D.shape) < 0.5] = 0.0
    D[:, 11] = 0.0                                                              # an empty cell: NaN column, as in R
    G, C = D.shape
    p, i, x = _csc(D)
    sums, n_pos, means = api.csc_gene_stats(p, i, x, G)
    assert np.array_equal(sums, D.sum(axis=1)) and np.array_equal(n_pos, ori.n_cells_expressing(D))
    assert np.array_equal(means, ori.row_means(D))
    for nf in (None, 1e5):
        Y, kept, chr_kept = ops.ingest_sparse_counts(p, i, x, G, oligo["chr_codes"], min_mean_expr_cutoff=1.0,
                                                     min_cells_per_gene=3, normalize_factor=nf)
        want, want_kept = ori.ingest_sparse_counts(p, i, x, G, 1.0, 3, nf)
        assert np.array_equal(kept, want_kept) and np.array_equal(chr_kept, oligo["chr_codes"][kept])
        assert Y.shape == want.shape and np.array_equal(np.isnan(Y), np.isnan(want)) and np.all(np.isnan(Y[:, 11]))
        ok = ~np.isnan(want)
        assert np.array_equal(Y[ok], want[ok])                                  # count data: bit-identical
    # no filter: the whole matrix, equal to the dense entry point
    Y, cs = api.csc_normalize(p, i, x, G, want_col_sums=True)
    assert np.array_equal(cs, D.sum(axis=0))
    cols = np.flatnonzero(cs > 0)
    dense = orc.normalize_by_seq_depth(D[:, cols], np.median(cs))
    assert np.array_equal(Y[:, cols], dense)


def test_scale_and_chromosome_end_removal(example_object):
    """run() step 5 (scale_data = TRUE) and step 13 (remove_genes_at_chr_ends = TRUE), both mirrored onto @.hspike."""
    from mirror import ops
    X = example_object["expr"]
    G, C = X.shape
    codes = example_object["chr_codes"]
    H = np.asfortranarray(X[:, :6] ** 2)
    obj = ops.Infercnv(expr_data=X, gene_order_chr=codes, gene_names=["g%d" % i for i in range(G)],
                       gene_order_start=np.arange(G), gene_order_stop=np.arange(G) + 5,
                       hspike=ops.Infercnv(expr_data=H, gene_order_chr=codes))
    out = ops.scale_infercnv_expr(obj)
    np.testing.assert_allclose(out.expr_data, ori.scale_infercnv_expr(X), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(out.hspike.expr_data, ori.scale_infercnv_expr(H), rtol=1e-11, atol=1e-13)
    const = X.copy(order="F")
    const[7] = 2.5
    assert np.all(np.isnan(ops.scale_infercnv_expr(ops.Infercnv(expr_data=const, gene_order_chr=codes)).expr_data[7]))   # 0 / 0
    cs, cl = orc.chr_ranges(codes)
    out = ops.remove_genes_at_ends_of_chromosomes(obj, 101)
    rm = ori.genes_removed_at_ends_of_chromosomes(cs, cl, 101)
    assert 0 < len(rm) < G
    keep = np.delete(np.arange(G), rm)
    assert np.array_equal(out.expr_data, X[keep]) and np.array_equal(out.hspike.expr_data, H[keep])
    assert out.gene_names == ["g%d" % i for i in keep] and np.array_equal(out.gene_order_chr, codes[keep])
    with pytest.raises(RuntimeError):
        ops.remove_genes_at_ends_of_chromosomes(obj, 5)           # tails below 3 genes: nothing to remove -> stop(1234)


def test_order_reduce_known_answers():
    """tests/testthat/test_infer_cnv.R:436-490."""
    from mirror import ops
    data = np.asfortranarray(np.tile(np.arange(1, 11, dtype=float)[:, None], (1, 2)))          # matrix(rep(1:10,2), ncol=2)
    names = ["gene_%d" % i for i in range(1, 11)]
    pos1 = ["gene_%d" % i for i in (10, 5, 8, 3, 4, 9, 1, 7, 6, 2)]
    r = ops.order_reduce(data, names, pos1, [1, 1, 2, 2, 3, 3, 4, 4, 5, 5], [1, 5] * 5, [4, 9] * 5)
    assert r["gene_names"] == pos1 and r["expr"][:, 0].tolist() == [10, 5, 8, 3, 4, 9, 1, 7, 6, 2]
    assert np.array_equal(r["expr"][:, 0], r["expr"][:, 1]) and r["chr"].tolist() == [1, 1, 2, 2, 3, 3, 4, 4, 5, 5]
    pos2 = ["gene_%d" % i for i in (10, 5, 3, 9, 1, 7)]                                          # dropping genes
    r = ops.order_reduce(data, names, pos2, [1, 1, 2, 3, 4, 4], [1, 5, 5, 5, 1, 5], [4, 9, 9, 9, 4, 9])
    assert r["gene_names"] == pos2 and r["expr"][:, 1].tolist() == [10, 5, 3, 9, 1, 7] and r["chr"].tolist() == [1, 1, 2, 3, 4, 4]
    pos3 = ["GENE_%d" % i for i in (10, 5, 3, 9, 1, 7)]                                          # no matching gene names
    assert ops.order_reduce(data, names, pos3, [1, 1, 2, 3, 4, 4], [1, 5, 5, 5, 1, 5], [4, 9, 9, 9, 4, 9])["expr"] is None
    assert ops.order_reduce(None, None, None, None, None, None)["expr"] is None
    # table not sorted, a chromosome name that sorts differently as text, an entry at position 0
    r = ops.order_reduce(data, names, ["gene_2", "gene_9", "gene_4", "gene_7", "gene_1"], ["chr10", "chr2", "chr10", "chr2", "chr2"],
                         [50, 7, 5, 0, 3], [60, 9, 8, 0, 4])
    assert r["gene_names"] == ["gene_4", "gene_2", "gene_1", "gene_9"] and r["expr"][:, 0].tolist() == [4, 2, 1, 9]
