"""The oracle's pairwise distances (oracle/infercnv_oracle.c: orc_pairwise_dist, the restatement of stats::dist's
R_euclidean / parallelDist's euclidean method in R's "dist" layout) against an independent implementation: the reference
holds no distance vectors of its own, scipy.spatial.distance.pdist is the pin (its condensed order - pairs (a, b), a < b, by
a then b - is exactly the strict lower triangle by columns that R stores)."""
import numpy as np
import pytest

from oracle import oracle as orc

pdist = pytest.importorskip("scipy.spatial.distance").pdist


@pytest.mark.parametrize("G,C", [(50, 7), (1, 3), (333, 40)])
def test_oracle_pairwise_dist_equals_scipy_pdist(G, C):
    rng = np.random.default_rng(G)
    X = np.asfortranarray(rng.normal(size=(G, C)))
    np.testing.assert_allclose(orc.pairwise_dist(X), pdist(X.T), rtol=1e-13, atol=0)
    cells = rng.permutation(C)[: max(2, C // 2)]
    np.testing.assert_allclose(orc.pairwise_dist(X, cells, nthreads=2), pdist(X[:, cells].T), rtol=1e-13, atol=0)


def test_oracle_pairwise_dist_layout_is_rs_dist_vector():
    """as.matrix(dist)[b, a] for a < b sits at n a - a (a + 1) / 2 + (b - a - 1) (0-based), R's documented layout"""
    X = np.asfortranarray(np.array([[0.0, 3.0, 0.0, 1.0], [0.0, 4.0, 1.0, 1.0]]))   # 2 genes x 4 cells
    d = orc.pairwise_dist(X)
    n = 4
    for a in range(n):
        for b in range(a + 1, n):
            assert d[n * a - a * (a + 1) // 2 + (b - a - 1)] == np.sqrt(np.sum((X[:, a] - X[:, b]) ** 2))
    assert d[0] == 5.0 and orc.pairwise_dist(X, [1]).size == 0
