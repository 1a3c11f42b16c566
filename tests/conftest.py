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
