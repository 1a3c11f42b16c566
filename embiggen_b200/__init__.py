"""B200-native Node2Vec / DeepWalk SkipGram & CBOW behind Embiggen's embedder API.

One hot path of monarch-initiative/embiggen (the Ensmallen-backed walk embedders,
/root/reference/embiggen/embedders/ensmallen_embedders/node2vec.py) rebuilt as hand-written
sm_100a CUDA behind a C ABI (``include/b2e.h``).  See DESIGN.md.
"""
from .graph import CSRGraph, as_csr, erdos_renyi, rmat, read_edge_list  # noqa: F401

__version__ = "0.1.0"
