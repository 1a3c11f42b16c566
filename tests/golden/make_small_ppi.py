"""Regenerates tests/golden/small_ppi_csr.npz from the reference's own fixture.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_small_ppi.py
The TSV (/root/reference/tests/data/small_ppi.tsv: 3 000 weighted, edge-typed undirected
edges) is read UNWEIGHTED and UNTYPED, node ids by sorted node name; this is BASELINE
config C1's graph (n = 1 064, nnz = 6 000).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from embiggen_b200.graph import read_edge_list  # noqa: E402

graph = read_edge_list("/root/reference/tests/data/small_ppi.tsv", name="small_ppi", weight_column=2)
assert graph.get_number_of_nodes() == 1064 and graph.indices.shape[0] == 6000
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "small_ppi_csr.npz"),
    indptr=graph.indptr, indices=graph.indices, node_names=np.array(graph.get_node_names()),
    weights=graph.weights,  # the fixture's native edge weights (the parity configs ignore them)
)
print("wrote small_ppi_csr.npz", graph.get_number_of_nodes(), graph.indices.shape[0])
