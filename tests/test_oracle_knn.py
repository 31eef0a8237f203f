"""The edge-builder oracle checks itself (CPU): the blocked shortlist k-NN equals the literal O(N^2) sort by
(distance, index), also with duplicated rows (exact ties), and the Pearson attribute is scipy.stats.pearsonr - the function
the reference calls (construct_graph/graph_constructor.py:278-280)."""
import numpy as np
import pytest
from scipy.stats import pearsonr

from oracle import knn as O


@pytest.mark.parametrize("n,f,radius,seed", [(60, 8, 6, 0), (257, 16, 9, 1), (33, 4, 33, 2), (1100, 12, 7, 3)])
def test_shortlist_knn_equals_literal_bruteforce(n, f, radius, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, f)).astype(np.float32)
    x[n // 2] = x[1]                                               # exact duplicates: ties broken by index
    x[n - 1] = x[1]
    a, b = O.exact_knn_edges(x, radius), O.exact_knn_edges_bruteforce(x, radius)
    assert a.dtype == np.int64 and a.shape == (2, n * (radius - 1))
    assert np.array_equal(a, b)
    assert np.array_equal(a[0], np.repeat(np.arange(n), radius - 1))           # i-major, radius-1 neighbours each


def test_too_few_nodes_raises_like_the_reference():
    with pytest.raises(ValueError):
        O.exact_knn_edges(np.zeros((4, 3), dtype=np.float32), 6)


def test_vectorised_pearson_equals_the_scipy_loop():
    """oracle.knn.pearson_edges (fp64, vectorised: what the large GPU cases are checked against) == the reference's
    per-edge scipy.stats.pearsonr loop (pearson_edges_scipy)"""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((40, 32)).astype(np.float32)
    ei = O.exact_knn_edges(x, 5)
    sim_v, et_v = O.pearson_edges(x, ei)
    sim_s, et_s = O.pearson_edges_scipy(x, ei)
    assert np.allclose(sim_v, sim_s, atol=1e-6) and np.array_equal(np.asarray(et_v), np.asarray(et_s))
    r = pearsonr(x[ei[0, 3]], x[ei[1, 3]])[0]
    assert abs(float(sim_s[3]) - float(r)) < 1e-12
