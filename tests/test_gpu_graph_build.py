"""GPU graph ingest (C ABI b2e_csr_from_edges / b2e_synthetic_csr) against the numpy builders."""
import numpy as np
import pytest

from embiggen_b200.graph import csr_from_edges, erdos_renyi, rmat, validate_csr
from embiggen_b200.graph_gpu import csr_from_edges_gpu, erdos_renyi_gpu, rmat_gpu

pytestmark = pytest.mark.gpu


def same_graph(a, b):
    return np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


@pytest.mark.parametrize("symmetrise", [True, False])
def test_csr_from_edges_matches_numpy(symmetrise):
    rng = np.random.default_rng(5)
    n = 5000
    src = rng.integers(0, n, 60000)
    dst = rng.integers(0, n, 60000)
    src[:100] = dst[:100]          # self-loops are dropped
    src[100:200] = src[200:300]    # duplicates are merged
    dst[100:200] = dst[200:300]
    expected = csr_from_edges(src, dst, n, symmetrise=symmetrise)
    got = csr_from_edges_gpu(src, dst, n, symmetrise=symmetrise)
    assert same_graph(expected, got)
    validate_csr(got.indptr, got.indices)
    assert got.is_directed() == (not symmetrise)


def test_csr_from_edges_small_ppi(small_ppi):
    """The reference's own fixture graph (tests/data/small_ppi.tsv) from its directed edge list."""
    src = np.repeat(np.arange(1064), np.diff(small_ppi.indptr))
    got = csr_from_edges_gpu(src, small_ppi.indices, 1064)
    assert same_graph(small_ppi, got)


def test_csr_from_edges_edge_cases():
    empty = csr_from_edges_gpu(np.zeros(0), np.zeros(0), 7)
    assert empty.indices.shape[0] == 0 and not empty.indptr.any()
    loops = csr_from_edges_gpu(np.arange(5), np.arange(5), 5)
    assert loops.indices.shape[0] == 0
    with pytest.raises(ValueError):
        csr_from_edges_gpu(np.array([0, 9]), np.array([1, 2]), 5)  # endpoint out of range


@pytest.mark.parametrize("n,m", [(2000, 12000), (300, 20000), (100_000, 1_000_000)])
def test_erdos_renyi_matches_numpy(n, m):
    assert same_graph(erdos_renyi(n, m, seed=7), erdos_renyi_gpu(n, m, seed=7))


@pytest.mark.parametrize("scale,m,n", [(12, 30000, 4000), (10, 5000, None), (17, 1_000_000, 100_000)])
def test_rmat_matches_numpy(scale, m, n):
    assert same_graph(rmat(scale, m, n=n, seed=11), rmat_gpu(scale, m, n=n, seed=11))


def test_full_size_shapes():
    """BASELINE C2 and C3 shapes: exact edge counts, simple, symmetric, sorted."""
    for graph, n, m in ((erdos_renyi_gpu(1_000_000, 10_000_000), 1_000_000, 10_000_000),
                        (rmat_gpu(24, 200_000_000, n=10_000_000), 10_000_000, 200_000_000)):
        assert graph.get_number_of_nodes() == n and graph.indices.shape[0] == 2 * m
        rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(graph.indptr))
        assert not (rows == graph.indices).any()                      # no self-loops
        order_ok = (graph.indices[1:] > graph.indices[:-1]) | (rows[1:] != rows[:-1])
        assert order_ok.all()                                         # sorted, no duplicates
        # symmetric: the multiset of (min, max) keys has every edge exactly twice
        lo = np.minimum(rows, graph.indices).astype(np.uint64)
        hi = np.maximum(rows, graph.indices).astype(np.uint64)
        keys = (lo << np.uint64(32)) | hi
        upper = keys[rows < graph.indices]
        lower = keys[rows > graph.indices]
        upper.sort()
        lower.sort()
        assert np.array_equal(upper, lower)
