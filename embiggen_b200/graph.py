"""CSR hand-off: the graph side of the drop-in boundary.

The reference passes an ``ensmallen.Graph`` to ``fit_transform``
(/root/reference/embiggen/utils/abstract_models/abstract_embedding_model.py:200-251) and
the only in-tree description of how a CSR is pulled out of it is
/root/reference/embiggen/embedders/pecanpy_embedders/node2vec.py:139-163
(``get_cumulative_node_degrees`` -> indptr[1:], ``get_directed_destination_node_ids`` ->
sorted neighbour arrays).  ``as_csr`` accepts such an object by duck typing, or a
:class:`CSRGraph`, a ``(indptr, indices)`` pair or a scipy sparse matrix, so the path runs
where the ``ensmallen`` wheel is absent.

Also here: seeded synthetic generators for the BASELINE.json shapes (Erdos-Renyi G(n, m),
R-MAT) and a reader for ``tests/data/small_ppi.tsv``-style edge lists.  They are host-side
numpy utilities for tests and benchmarks, not part of the hot path.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(seed: int, c0, c1, c2, c3) -> Tuple[np.ndarray, ...]:
    """Vectorised Philox4x32-10; counters are broadcastable uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    )
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> _S32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> _S32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


class CSRGraph:
    """Minimal stand-in for ``ensmallen.Graph`` exposing the accessors this path calls."""

    def __init__(self, indptr, indices, node_names: Optional[Sequence[str]] = None,
                 weights=None, name: str = "graph", directed: bool = False,
                 node_types=None, edge_types=None):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32)
        if self.indptr.ndim != 1 or self.indptr.shape[0] < 1:
            raise ValueError("indptr must be a 1-d array of length n + 1.")
        if self.indptr[0] != 0 or self.indptr[-1] != self.indices.shape[0]:
            raise ValueError("indptr must start at 0 and end at len(indices).")
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        self.node_types = None if node_types is None else np.ascontiguousarray(node_types, dtype=np.uint32)
        self.edge_types = None if edge_types is None else np.ascontiguousarray(edge_types, dtype=np.uint32)
        if self.node_types is not None and self.node_types.shape != (self.indptr.shape[0] - 1,):
            raise ValueError("node_types must have one entry per node.")
        if self.edge_types is not None and self.edge_types.shape != self.indices.shape:
            raise ValueError("edge_types must have one entry per directed edge.")
        self._node_names = None if node_names is None else list(node_names)
        self._name = name
        self._directed = directed

    # -- accessors mirrored from ensmallen.Graph (call sites cited in the module docstring) --
    def get_name(self) -> str:
        return self._name

    def get_number_of_nodes(self) -> int:
        return int(self.indptr.shape[0] - 1)

    def get_number_of_directed_edges(self) -> int:
        return int(self.indices.shape[0])

    def get_cumulative_node_degrees(self) -> np.ndarray:
        return self.indptr[1:].astype(np.uint64)

    def get_directed_destination_node_ids(self) -> np.ndarray:
        return self.indices

    def get_directed_edge_weights(self) -> np.ndarray:
        if self.weights is None:
            raise ValueError("The graph has no edge weights.")
        return self.weights

    def get_node_names(self) -> List[str]:
        if self._node_names is None:
            return [str(i) for i in range(self.get_number_of_nodes())]
        return self._node_names

    def get_node_degrees(self) -> np.ndarray:
        return np.diff(self.indptr).astype(np.uint32)

    def has_nodes(self) -> bool:
        return self.get_number_of_nodes() > 0

    def has_edges(self) -> bool:
        return self.indices.shape[0] > 0

    def has_edge_weights(self) -> bool:
        return self.weights is not None

    def has_negative_edge_weights(self) -> bool:
        return self.weights is not None and bool((self.weights < 0).any())

    def has_node_types(self) -> bool:
        return self.node_types is not None

    def has_edge_types(self) -> bool:
        return self.edge_types is not None

    def get_number_of_node_types(self) -> int:
        return 0 if self.node_types is None else int(np.unique(self.node_types).shape[0])

    def get_number_of_edge_types(self) -> int:
        return 0 if self.edge_types is None else int(np.unique(self.edge_types).shape[0])

    def get_single_label_node_type_ids(self) -> np.ndarray:
        if self.node_types is None:
            raise ValueError("The graph has no node types.")
        return self.node_types

    def get_directed_edge_type_ids(self) -> np.ndarray:
        if self.edge_types is None:
            raise ValueError("The graph has no edge types.")
        return self.edge_types

    def is_directed(self) -> bool:
        return self._directed

    def get_number_of_disconnected_nodes(self) -> int:
        return int((np.diff(self.indptr) == 0).sum())

    def has_disconnected_nodes(self) -> bool:
        return self.get_number_of_disconnected_nodes() > 0

    def has_nodes_sorted_by_decreasing_outbound_node_degree(self) -> bool:
        degrees = np.diff(self.indptr)
        return bool((degrees[:-1] >= degrees[1:]).all())


def validate_csr(indptr: np.ndarray, indices: np.ndarray) -> None:
    """Raise ValueError unless rows are sorted ascending and ids are in range."""
    n = indptr.shape[0] - 1
    if (np.diff(indptr) < 0).any():
        raise ValueError("indptr must be non-decreasing.")
    if indices.shape[0] and int(indices.max()) >= n:
        raise ValueError("A destination node id is out of range.")
    if indices.shape[0] > 1:
        unsorted = indices[1:] < indices[:-1]
        row_starts = indptr[1:-1]
        row_starts = row_starts[(row_starts > 0) & (row_starts < indices.shape[0])]
        unsorted[row_starts - 1] = False
        if unsorted.any():
            raise ValueError("Neighbour lists must be sorted ascending within each row.")


def as_csr(graph) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """Return (indptr int64[n+1], indices uint32[nnz], weights float32[nnz] or None)."""
    if isinstance(graph, CSRGraph):
        return graph.indptr, graph.indices, graph.weights
    if isinstance(graph, (tuple, list)) and len(graph) in (2, 3):
        g = CSRGraph(*graph[:2], weights=graph[2] if len(graph) == 3 else None)
        return g.indptr, g.indices, g.weights
    if hasattr(graph, "indptr") and hasattr(graph, "indices") and hasattr(graph, "tocsr"):
        m = graph.tocsr()
        m.sort_indices()
        return (np.ascontiguousarray(m.indptr, dtype=np.int64),
                np.ascontiguousarray(m.indices, dtype=np.uint32), None)
    if hasattr(graph, "get_cumulative_node_degrees"):
        n = int(graph.get_number_of_nodes())
        indptr = np.zeros(n + 1, dtype=np.int64)
        indptr[1:] = np.asarray(graph.get_cumulative_node_degrees(), dtype=np.int64)
        indices = np.ascontiguousarray(graph.get_directed_destination_node_ids(), dtype=np.uint32)
        weights = None
        if graph.has_edge_weights():
            weights = np.ascontiguousarray(graph.get_directed_edge_weights(), dtype=np.float32)
        return indptr, indices, weights
    raise ValueError(
        f"Cannot extract a CSR from an object of type {type(graph)}: expected an "
        "ensmallen.Graph, a CSRGraph, a (indptr, indices) pair or a scipy sparse matrix."
    )


def as_types(graph) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """(node type id per node, edge type id per directed edge in CSR order), each None when the
    graph has none; read through the accessors the reference uses on an ``ensmallen.Graph``
    (``get_single_label_node_type_ids``, ``get_directed_edge_type_ids``)."""
    node_types = edge_types = None
    if hasattr(graph, "has_node_types") and graph.has_node_types():
        node_types = np.ascontiguousarray(graph.get_single_label_node_type_ids(), dtype=np.uint32)
    if hasattr(graph, "has_edge_types") and graph.has_edge_types():
        edge_types = np.ascontiguousarray(graph.get_directed_edge_type_ids(), dtype=np.uint32)
    return node_types, edge_types


def as_graph(graph):
    """Anything ``as_csr`` accepts, as an object with the ``ensmallen.Graph`` accessors."""
    if hasattr(graph, "get_number_of_nodes") and hasattr(graph, "has_edges"):
        return graph
    indptr, indices, weights = as_csr(graph)
    return CSRGraph(indptr, indices, weights=weights)


def csr_from_edges(src: np.ndarray, dst: np.ndarray, n: int, symmetrise: bool = True,
                   node_names=None, name: str = "graph", weights=None) -> CSRGraph:
    """Sorted, de-duplicated, self-loop-free CSR from an edge list (a duplicated edge keeps the
    weight of its first occurrence)."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if weights is not None:
        weights = np.asarray(weights, dtype=np.float32)[keep]
    if symmetrise:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
        if weights is not None:
            weights = np.concatenate([weights, weights])
    keys, first = np.unique(src * np.int64(n) + dst, return_index=True)
    rows = keys // np.int64(n)
    indices = (keys - rows * np.int64(n)).astype(np.uint32)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    return CSRGraph(indptr, indices, node_names=node_names, name=name, directed=not symmetrise,
                    weights=None if weights is None else weights[first])


def erdos_renyi(n: int, m: int, seed: int = 42) -> CSRGraph:
    """G(n, m): m distinct undirected edges, no self-loops (BASELINE config C2)."""
    if m > n * (n - 1) // 2:
        raise ValueError("Too many edges requested.")
    chosen = np.empty(0, dtype=np.int64)
    drawn = 0
    while chosen.shape[0] < m:
        want = int((m - chosen.shape[0]) * 1.1) + 16
        idx = np.arange(drawn, drawn + want, dtype=np.uint64)
        drawn += want
        r0, r1, _, _ = philox4x32(seed, idx & _MASK, idx >> _S32, 0, 0x10 << 24)
        u = ((r0.astype(np.uint64) * np.uint64(n)) >> _S32).astype(np.int64)
        v = ((r1.astype(np.uint64) * np.uint64(n)) >> _S32).astype(np.int64)
        ok = u != v
        lo, hi = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
        keys = np.concatenate([chosen, lo * np.int64(n) + hi])
        _, first = np.unique(keys, return_index=True)
        first.sort()
        chosen = keys[first][:m]
    return csr_from_edges(chosen // n, chosen % n, n, name=f"ER_{n}_{m}")


def rmat(scale: int, m: int, n: Optional[int] = None, seed: int = 42,
         probabilities=(0.57, 0.19, 0.19, 0.05)) -> CSRGraph:
    """R-MAT with ids >= n rejected; m distinct undirected edges after dedup (C3-C5)."""
    n = (1 << scale) if n is None else n
    a, b, c, _ = probabilities
    t_a = np.uint64(int(a * 2 ** 32))
    t_ab = np.uint64(int((a + b) * 2 ** 32))
    t_abc = np.uint64(int((a + b + c) * 2 ** 32))
    chosen = np.empty(0, dtype=np.int64)
    drawn = 0
    while chosen.shape[0] < m:
        want = int((m - chosen.shape[0]) * 1.25) + 16
        idx = np.arange(drawn, drawn + want, dtype=np.uint64)
        drawn += want
        u = np.zeros(want, dtype=np.int64)
        v = np.zeros(want, dtype=np.int64)
        for block in range((scale + 3) // 4):
            words = philox4x32(seed, idx & _MASK, idx >> _S32, block, 0x11 << 24)
            for level in range(4):
                if block * 4 + level >= scale:
                    break
                r = words[level].astype(np.uint64)
                bit_u = (r >= t_ab).astype(np.int64)
                bit_v = (((r >= t_a) & (r < t_ab)) | (r >= t_abc)).astype(np.int64)
                u = (u << 1) | bit_u
                v = (v << 1) | bit_v
        ok = (u != v) & (u < n) & (v < n)
        lo, hi = np.minimum(u[ok], v[ok]), np.maximum(u[ok], v[ok])
        keys = np.concatenate([chosen, lo * np.int64(n) + hi])
        _, first = np.unique(keys, return_index=True)
        first.sort()
        chosen = keys[first][:m]
    return csr_from_edges(chosen // n, chosen % n, n, name=f"RMAT_{scale}_{m}")


def read_edge_list(path: str, source_column: int = 0, destination_column: int = 1,
                   header: bool = True, separator: str = "\t", name: Optional[str] = None,
                   weight_column: Optional[int] = None) -> CSRGraph:
    """Undirected graph from a TSV edge list (e.g. tests/data/small_ppi.tsv), weighted when a
    weight column is named.

    Node ids are assigned by sorted node name, as GRAPE does for an unsorted vocabulary.
    """
    sources, destinations, weights = [], [], []
    with open(path) as handle:
        if header:
            next(handle)
        for line in handle:
            fields = line.rstrip("\n").split(separator)
            if len(fields) <= max(source_column, destination_column, weight_column or 0):
                continue
            sources.append(fields[source_column])
            destinations.append(fields[destination_column])
            if weight_column is not None:
                weights.append(float(fields[weight_column]))
    names = sorted(set(sources) | set(destinations))
    ids = {node: i for i, node in enumerate(names)}
    src = np.fromiter((ids[s] for s in sources), dtype=np.int64, count=len(sources))
    dst = np.fromiter((ids[d] for d in destinations), dtype=np.int64, count=len(destinations))
    return csr_from_edges(src, dst, len(names), node_names=names, name=name or path,
                          weights=weights if weight_column is not None else None)
