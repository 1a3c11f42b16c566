"""Graph ingest on the GPU through the C ABI.

``csr_from_edges_gpu`` builds the sorted, de-duplicated CSR the reference receives from an
``ensmallen.Graph`` (/root/reference/embiggen/embedders/pecanpy_embedders/node2vec.py:139-163)
out of a plain edge list; ``erdos_renyi_gpu`` / ``rmat_gpu`` generate the synthetic graphs of the
BASELINE.json shapes and return the same graphs as the numpy generators of
:mod:`embiggen_b200.graph` (same Philox stream, same "first m distinct edges in draw order"
rule), which is how the 200 M-edge configurations become practical.
"""
import ctypes
from typing import Optional

import numpy as np

from . import _lib
from ._lib import check
from .graph import CSRGraph


def csr_from_edges_gpu(src, dst, n: int, symmetrise: bool = True, node_names=None,
                       name: str = "graph", device: int = 0) -> CSRGraph:
    src = np.ascontiguousarray(src, dtype=np.uint32)
    dst = np.ascontiguousarray(dst, dtype=np.uint32)
    if src.shape != dst.shape or src.ndim != 1:
        raise ValueError("src and dst must be 1-d arrays of the same length.")
    capacity = max(1, src.shape[0] * (2 if symmetrise else 1))
    indptr = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(capacity, dtype=np.uint32)
    nnz = ctypes.c_uint64()
    check(_lib.load().b2e_csr_from_edges(device, src.ctypes.data, dst.ctypes.data, src.shape[0], n,
                                         int(symmetrise), indptr.ctypes.data, indices.ctypes.data,
                                         capacity, ctypes.byref(nnz)))
    return CSRGraph(indptr, indices[: nnz.value].copy(), node_names=node_names, name=name,
                    directed=not symmetrise)


def _synthetic(kind: int, n: int, scale: int, m: int, seed: int, thresholds, name: str,
               device: int) -> CSRGraph:
    indptr = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(2 * m, dtype=np.uint32)
    nnz = ctypes.c_uint64()
    check(_lib.load().b2e_synthetic_csr(device, kind, n, scale, m, seed, *thresholds,
                                        indptr.ctypes.data, indices.ctypes.data, 2 * m,
                                        ctypes.byref(nnz)))
    assert nnz.value == 2 * m
    return CSRGraph(indptr, indices, name=name)


def erdos_renyi_gpu(n: int, m: int, seed: int = 42, device: int = 0) -> CSRGraph:
    """G(n, m), identical to :func:`embiggen_b200.graph.erdos_renyi`."""
    return _synthetic(0, n, 0, m, seed, (0, 0, 0), f"ER_{n}_{m}", device)


def rmat_gpu(scale: int, m: int, n: Optional[int] = None, seed: int = 42,
             probabilities=(0.57, 0.19, 0.19, 0.05), device: int = 0) -> CSRGraph:
    """R-MAT, identical to :func:`embiggen_b200.graph.rmat`."""
    n = (1 << scale) if n is None else n
    a, b, c, _ = probabilities
    thresholds = (int(a * 2 ** 32), int((a + b) * 2 ** 32), int((a + b + c) * 2 ** 32))
    return _synthetic(1, n, scale, m, seed, thresholds, f"RMAT_{scale}_{m}", device)
