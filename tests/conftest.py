import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def small_ppi():
    """BASELINE config C1's graph (derived from the reference fixture, see make_small_ppi.py)."""
    from embiggen_b200.graph import CSRGraph
    data = np.load(os.path.join(GOLDEN, "small_ppi_csr.npz"))
    return CSRGraph(data["indptr"], data["indices"], node_names=list(data["node_names"]),
                    name="small_ppi")


@pytest.fixture(scope="session")
def small_ppi_weighted():
    """The same graph with the fixture's native edge weights (700 .. 999)."""
    from embiggen_b200.graph import CSRGraph
    data = np.load(os.path.join(GOLDEN, "small_ppi_csr.npz"))
    return CSRGraph(data["indptr"], data["indices"], node_names=list(data["node_names"]),
                    weights=data["weights"], name="small_ppi_weighted")


@pytest.fixture(scope="session")
def er_graph():
    from embiggen_b200.graph import erdos_renyi
    return erdos_renyi(2000, 12000, seed=7)


@pytest.fixture(scope="session")
def rmat_graph():
    from embiggen_b200.graph import rmat
    return rmat(12, 30000, n=4000, seed=11)


def tiny_graphs():
    """Edge cases: path, star (hub), triangle + pendant, two components, isolated node."""
    from embiggen_b200.graph import csr_from_edges
    graphs = {}
    graphs["path"] = csr_from_edges(np.arange(9), np.arange(1, 10), 10, name="path")
    graphs["star"] = csr_from_edges(np.zeros(40, dtype=np.int64), np.arange(1, 41), 41, name="star")
    graphs["triangle_pendant"] = csr_from_edges(
        np.array([0, 1, 2, 2]), np.array([1, 2, 0, 3]), 4, name="triangle_pendant")
    graphs["two_components_isolated"] = csr_from_edges(
        np.array([0, 1, 4, 5, 6]), np.array([1, 2, 5, 6, 4]), 8, name="two_components_isolated")
    graphs["directed_dead_end"] = csr_from_edges(
        np.array([0, 1, 2, 0]), np.array([1, 2, 3, 2]), 5, symmetrise=False,
        name="directed_dead_end")
    return graphs


def heldout_sgns_loss(graph, central, contextual, window=4, negatives=5, n_walks=400, length=32,
                      seed=987654321, return_weight=1.0, explore_weight=1.0):
    """Mean SGNS objective of (central, contextual) tables on walks neither run trained on.

    The same fixed sample of (centre, context, negatives) is scored for every embedding handed
    in, which is how two training runs are compared (leg 3)."""
    import oracle
    walks, _ = oracle.walks(graph.indptr, graph.indices, seed, 10 ** 9, n_walks, length,
                            return_weight, explore_weight)
    rng = np.random.default_rng(seed)
    n = graph.get_number_of_nodes()
    total, count = 0.0, 0
    c64, x64 = central.astype(np.float64), contextual.astype(np.float64)
    for offset in range(1, window + 1):
        a, b = walks[:, :-offset].ravel(), walks[:, offset:].ravel()
        keep = (a != oracle.PAD_TOKEN) & (b != oracle.PAD_TOKEN) & (a != b)
        a, b = a[keep].astype(np.int64), b[keep].astype(np.int64)
        for centre, context in ((a, b), (b, a)):
            positive = np.einsum("ij,ij->i", c64[centre], x64[context])
            total += np.logaddexp(0.0, -positive).sum()
            neg = rng.integers(0, n, size=(centre.shape[0], negatives))
            scores = np.einsum("ij,ikj->ik", c64[centre], x64[neg])
            total += np.logaddexp(0.0, scores).sum()
            count += centre.shape[0]
    return total / count
