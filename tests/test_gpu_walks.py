"""GPU walks (through the C ABI) must equal the CPU oracle's walks bit for bit."""
import numpy as np
import pytest

import oracle
from conftest import tiny_graphs
from embiggen_b200.engine import Engine

pytestmark = pytest.mark.gpu

PQ = [(1.0, 1.0), (0.25, 4.0), (2.0, 0.5), (0.5, 2.0), (1.0, 3.0), (7.5, 1.0)]


def gpu_walks(graph, seed, first, count, length, rw, ew, stride=1, chunk=0):
    with Engine("SkipGram", walk_length=length, return_weight=rw, explore_weight=ew,
                iterations=1, chunk_walks=chunk) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        walks = engine.walks(seed, first, count, stride)
        counters = engine.counters()
    return walks, counters


@pytest.mark.parametrize("rw,ew", PQ)
@pytest.mark.parametrize("fixture", ["small_ppi", "er_graph", "rmat_graph"])
def test_walks_bit_exact(request, fixture, rw, ew):
    graph = request.getfixturevalue(fixture)
    n_src = int((np.diff(graph.indptr) > 0).sum())
    count = 2 * n_src + 17  # more than one iteration, ragged tail
    expected, oc = oracle.walks(graph.indptr, graph.indices, 42, 5, count, 128, rw, ew)
    got, gc = gpu_walks(graph, 42, 5, count, 128, rw, ew)
    assert np.array_equal(got, expected)
    assert gc["walk_steps"] == oc["steps"]
    assert gc["walk_trials"] == oc["trials"]
    assert gc["walk_searches"] == oc["searches"]


@pytest.mark.parametrize("length", [2, 3, 4, 5, 7, 33, 130])
def test_ragged_walk_lengths(small_ppi, length):
    expected, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 9, 0, 777, length, 0.25, 4.0)
    got, _ = gpu_walks(small_ppi, 9, 0, 777, length, 0.25, 4.0)
    assert np.array_equal(got, expected)


@pytest.mark.parametrize("name", sorted(tiny_graphs()))
@pytest.mark.parametrize("rw,ew", [(1.0, 1.0), (0.25, 4.0), (4.0, 0.25)])
def test_edge_case_graphs(name, rw, ew):
    graph = tiny_graphs()[name]
    expected, _ = oracle.walks(graph.indptr, graph.indices, 3, 0, 64, 16, rw, ew)
    got, _ = gpu_walks(graph, 3, 0, 64, 16, rw, ew)
    assert np.array_equal(got, expected)
    if name == "directed_dead_end":
        assert (got == oracle.PAD_TOKEN).any()


def test_walk_ids_shard_invariant(er_graph):
    """Counter-based RNG: a walk id gives the same walk whatever the sharding / chunking."""
    whole, _ = gpu_walks(er_graph, 42, 0, 4000, 64, 0.5, 2.0)
    for world in (2, 4, 8):
        for rank in range(world):
            count = (4000 - rank + world - 1) // world
            shard, _ = gpu_walks(er_graph, 42, rank, count, 64, 0.5, 2.0, stride=world, chunk=257)
            assert np.array_equal(shard, whole[rank::world])


def test_seed_and_64bit_walk_ids(er_graph):
    first = (1 << 40) + 12345
    seed = 0xDEADBEEFCAFEF00D
    expected, _ = oracle.walks(er_graph.indptr, er_graph.indices, seed, first, 500, 32, 0.25, 4.0)
    got, _ = gpu_walks(er_graph, seed, first, 500, 32, 0.25, 4.0)
    assert np.array_equal(got, expected)
    other, _ = gpu_walks(er_graph, seed + 1, first, 500, 32, 0.25, 4.0)
    assert not np.array_equal(got, other)


def test_full_size_properties():
    """BASELINE-scale shape (ER 1M / 10M): every transition is an edge; start nodes cycle."""
    from embiggen_b200.graph import erdos_renyi
    graph = erdos_renyi(1_000_000, 10_000_000, seed=42)
    n_src = int((np.diff(graph.indptr) > 0).sum())
    count = 200_000
    walks, counters = gpu_walks(graph, 42, n_src - 1000, count, 128, 1.0, 1.0)
    assert counters["walk_steps"] == count * 127
    sources = np.flatnonzero(np.diff(graph.indptr) > 0)
    ids = (n_src - 1000 + np.arange(count)) % n_src
    assert np.array_equal(walks[:, 0], sources[ids].astype(np.uint32))
    src = walks[:, :-1].ravel().astype(np.int64)
    dst = walks[:, 1:].ravel().astype(np.int64)
    keys = src * graph.get_number_of_nodes() + dst
    edge_keys = (np.repeat(np.arange(graph.get_number_of_nodes(), dtype=np.int64),
                           np.diff(graph.indptr)) * graph.get_number_of_nodes()
                 + graph.indices.astype(np.int64))
    assert np.isin(keys[:2_000_000], edge_keys).all()
    # the oracle agrees on a bounded sample of the same walks
    expected, _ = oracle.walks(graph.indptr, graph.indices, 42, n_src - 1000, 2000, 128)
    assert np.array_equal(walks[:2000], expected)


# ---- edge weights (C ABI b2e_load_csr_weighted) ----
def gpu_weighted_walks(graph, weights, seed, first, count, length, rw, ew):
    with Engine("SkipGram", walk_length=length, return_weight=rw, explore_weight=ew,
                iterations=1) as engine:
        engine.load_csr(graph.indptr, graph.indices, weights)
        return engine.walks(seed, first, count), engine.counters()


@pytest.mark.parametrize("rw,ew", [(1.0, 1.0), (0.25, 4.0), (2.0, 0.5)])
def test_weighted_walks_bit_exact(small_ppi_weighted, er_graph, rw, ew):
    g = small_ppi_weighted
    expected, oc = oracle.walks(g.indptr, g.indices, 42, 3, 2500, 64, rw, ew, weights=g.weights)
    got, gc = gpu_weighted_walks(g, g.weights, 42, 3, 2500, 64, rw, ew)
    assert np.array_equal(got, expected)
    assert (gc["walk_steps"], gc["walk_trials"], gc["walk_searches"]) == (oc["steps"], oc["trials"], oc["searches"])
    # random weights over six orders of magnitude, some exactly zero
    rng = np.random.default_rng(1)
    weights = np.exp(rng.uniform(-7, 7, er_graph.indices.shape[0])).astype(np.float32)
    weights[rng.random(weights.shape[0]) < 0.05] = 0.0
    expected, _ = oracle.walks(er_graph.indptr, er_graph.indices, 9, 0, 3000, 40, rw, ew, weights=weights)
    got, _ = gpu_weighted_walks(er_graph, weights, 9, 0, 3000, 40, rw, ew)
    assert np.array_equal(got, expected)


def test_weighted_embedder_runs_and_rejects_negative_weights(small_ppi_weighted):
    from embiggen_b200.embedders import Node2VecSkipGramB200
    from embiggen_b200.graph import CSRGraph
    model = Node2VecSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1, verbose=False)
    tables = model.fit_transform(small_ppi_weighted, return_dataframe=False).get_all_node_embedding()
    assert np.isfinite(tables[0]).all()
    g = small_ppi_weighted
    with Engine("SkipGram") as engine:
        with pytest.raises(ValueError):
            engine.load_csr(g.indptr, g.indices, -g.weights)
    negative = CSRGraph(g.indptr, g.indices, weights=-g.weights, name="negative")
    with pytest.raises(ValueError, match="negative edge weights"):
        model.fit_transform(negative)


def test_fold_filter_and_short_row_variants(monkeypatch, small_ppi, rmat_graph):
    """The folded return edge is part of the specification (oracle/walks.c): an undirected graph
    with return_weight > max(1, explore_weight) walks on Philox tag 12, anything else on tag 2.
    The row filters and the short-row search are accelerators: switching them off must not
    change a token or a decision counter, only the number of probes."""
    graphs = [small_ppi, rmat_graph, tiny_graphs()["directed_dead_end"], tiny_graphs()["star"]]
    for graph in graphs:
        for rw, ew in [(0.25, 4.0), (2.0, 0.5), (7.5, 1.0)]:
            for length in (2, 5, 64, 130):
                count = 3 * int((np.diff(graph.indptr) > 0).sum()) + 5
                expected, oc = oracle.walks(graph.indptr, graph.indices, 42, 9, count, length, rw, ew)
                got, gc = gpu_walks(graph, 42, 9, count, length, rw, ew)
                assert np.array_equal(got, expected)
                assert (gc["walk_steps"], gc["walk_trials"], gc["walk_searches"]) == \
                    (oc["steps"], oc["trials"], oc["searches"])
    # the fold halves the trials of C3's p/q on an undirected graph and is a different stream
    folded, fc = oracle.walks(rmat_graph.indptr, rmat_graph.indices, 1, 0, 5000, 64, 2.0, 0.5)
    plain, pc = oracle.walks(rmat_graph.indptr, rmat_graph.indices, 1, 0, 5000, 64, 2.0, 0.5, undirected=False)
    assert not np.array_equal(folded, plain) and fc["trials"] < 0.7 * pc["trials"]
    base, bc = gpu_walks(rmat_graph, 1, 0, 5000, 64, 2.0, 0.5)
    assert np.array_equal(base, folded) and bc["walk_filter_rejects"] > 0
    monkeypatch.setenv("B2E_NO_FILTER", "1")
    got, gc = gpu_walks(rmat_graph, 1, 0, 5000, 64, 2.0, 0.5)
    assert np.array_equal(got, folded) and gc["walk_filter_rejects"] == 0
    assert gc["walk_probes"] > bc["walk_probes"] and gc["walk_searches"] == bc["walk_searches"]
    monkeypatch.delenv("B2E_NO_FILTER")
    monkeypatch.setenv("B2E_NO_FOLD", "1")
    got, gc = gpu_walks(rmat_graph, 1, 0, 5000, 64, 2.0, 0.5)
    assert np.array_equal(got, plain) and gc["walk_trials"] == pc["trials"]
    monkeypatch.delenv("B2E_NO_FOLD")
    monkeypatch.setenv("B2E_ASSUME_DIRECTED", "1")  # no fold, no short-row search
    got, _ = gpu_walks(rmat_graph, 1, 0, 5000, 64, 2.0, 0.5)
    assert np.array_equal(got, plain)


def test_load_rejects_malformed_csr_and_leaves_a_clean_handle(er_graph):
    """b2e_load_csr checks the CSR contents on the device: ids in range, rows strictly ascending
    (the reference's graph object guarantees both; a raw (indptr, indices) pair does not)."""
    indptr, indices = er_graph.indptr, er_graph.indices.copy()
    row = int(np.flatnonzero(np.diff(indptr) >= 3)[0])
    swapped = indices.copy()
    swapped[indptr[row]], swapped[indptr[row] + 1] = indices[indptr[row] + 1], indices[indptr[row]]
    duplicate = indices.copy()
    duplicate[indptr[row] + 1] = duplicate[indptr[row]]
    out_of_range = indices.copy()
    out_of_range[-1] = er_graph.get_number_of_nodes()
    with Engine("SkipGram", return_weight=2.0, explore_weight=0.5) as engine:
        for bad, message in ((swapped, "sorted"), (duplicate, "sorted"), (out_of_range, "out of range")):
            with pytest.raises(ValueError, match=message):
                engine.load_csr(indptr, bad)
            with pytest.raises(RuntimeError, match="b2e_load_csr"):
                engine.init_tables(1)
        engine.load_csr(indptr, indices)  # the handle is still usable
        assert engine.walks(1, 0, 10).shape == (10, 128)


# ---- normalize_by_degree (b2e_config.normalize_by_degree; walk_norm_kernel) ----
@pytest.mark.parametrize("rw,ew", [(1.0, 1.0), (0.25, 4.0), (2.0, 0.5)])
def test_normalize_by_degree_walks_bit_exact(small_ppi_weighted, rmat_graph, rw, ew):
    from conftest import tiny_graphs
    cases = [(rmat_graph, None, 64), (small_ppi_weighted, small_ppi_weighted.weights, 33),
             (small_ppi_weighted, None, 130), (tiny_graphs()["directed_dead_end"], None, 9),
             (tiny_graphs()["star"], None, 8)]
    for graph, weights, length in cases:
        count = 2 * int((np.diff(graph.indptr) > 0).sum()) + 3
        expected, oc = oracle.walks(graph.indptr, graph.indices, 42, 5, count, length, rw, ew,
                                    weights=weights, normalize_by_degree=True)
        with Engine("SkipGram", walk_length=length, return_weight=rw, explore_weight=ew,
                    iterations=1, normalize_by_degree=True) as engine:
            engine.load_csr(graph.indptr, graph.indices, weights)
            got, gc = engine.walks(42, 5, count), engine.counters()
        assert np.array_equal(got, expected)
        assert (gc["walk_steps"], gc["walk_trials"], gc["walk_searches"]) == \
            (oc["steps"], oc["trials"], oc["searches"])
        assert oc["capped"] == 0


def test_normalize_by_degree_embedder_runs(small_ppi):
    from embiggen_b200.embedders import Node2VecCBOWB200
    model = Node2VecCBOWB200(embedding_size=8, epochs=1, walk_length=8, iterations=1,
                             normalize_by_degree=True, verbose=False)
    tables = model.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    assert all(np.isfinite(t).all() for t in tables)


# ---- typed walks (b2e_load_types, change_node_type_weight / change_edge_type_weight) ----
def _types_for(graph, seed):
    rng = np.random.default_rng(seed)
    n = graph.get_number_of_nodes()
    node_types = rng.integers(0, 4, n).astype(np.uint32)
    rows = np.repeat(np.arange(n), np.diff(graph.indptr))
    lo, hi = np.minimum(rows, graph.indices), np.maximum(rows, graph.indices)
    edge_types = ((lo * 2654435761 + hi * 40503) % 3).astype(np.uint32)  # one type per undirected edge
    return node_types, edge_types


@pytest.mark.parametrize("rw,ew,cn,ce", [(1.0, 1.0, 3.0, 1.0), (1.0, 1.0, 1.0, 0.25),
                                         (0.25, 4.0, 0.2, 5.0), (2.0, 0.5, 4.0, 0.5)])
def test_typed_walks_bit_exact(small_ppi_weighted, rmat_graph, rw, ew, cn, ce):
    from conftest import tiny_graphs
    cases = [(rmat_graph, None, 64, False), (small_ppi_weighted, small_ppi_weighted.weights, 33, False),
             (small_ppi_weighted, None, 130, True), (tiny_graphs()["directed_dead_end"], None, 9, False),
             (tiny_graphs()["star"], None, 8, True)]
    for graph, weights, length, normalize in cases:
        node_types, edge_types = _types_for(graph, 3)
        count = 2 * int((np.diff(graph.indptr) > 0).sum()) + 3
        expected, oc = oracle.walks(graph.indptr, graph.indices, 42, 5, count, length, rw, ew,
                                    weights=weights, normalize_by_degree=normalize,
                                    node_types=node_types, edge_types=edge_types,
                                    change_node_type_weight=cn, change_edge_type_weight=ce)
        with Engine("SkipGram", walk_length=length, return_weight=rw, explore_weight=ew,
                    iterations=1, normalize_by_degree=normalize, change_node_type_weight=cn,
                    change_edge_type_weight=ce) as engine:
            engine.load_csr(graph.indptr, graph.indices, weights)
            engine.load_types(node_types, edge_types)
            got, gc = engine.walks(42, 5, count), engine.counters()
        assert np.array_equal(got, expected)
        assert (gc["walk_steps"], gc["walk_trials"], gc["walk_searches"]) == \
            (oc["steps"], oc["trials"], oc["searches"])
        assert oc["capped"] == 0


def test_typed_walks_without_types_or_with_unit_weights_are_plain(small_ppi):
    node_types, edge_types = _types_for(small_ppi, 1)
    plain = gpu_walks(small_ppi, 7, 0, 2000, 40, 0.25, 4.0)[0]
    with Engine("SkipGram", walk_length=40, return_weight=0.25, explore_weight=4.0, iterations=1,
                change_node_type_weight=3.0) as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)  # no types loaded: untyped walks
        assert np.array_equal(engine.walks(7, 0, 2000), plain)
        engine.load_types(node_types, None)
        assert not np.array_equal(engine.walks(7, 0, 2000), plain)
        engine.load_csr(small_ppi.indptr, small_ppi.indices)  # a new graph drops the types
        assert np.array_equal(engine.walks(7, 0, 2000), plain)
    with Engine("SkipGram", walk_length=40, return_weight=0.25, explore_weight=4.0, iterations=1) as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        engine.load_types(node_types, edge_types)  # unit change weights: plain kernel
        assert np.array_equal(engine.walks(7, 0, 2000), plain)
        with pytest.raises(ValueError):
            engine.load_types(node_types[:-1], None)


def test_typed_embedder(small_ppi):
    from embiggen_b200.embedders import Node2VecSkipGramB200
    from embiggen_b200.graph import CSRGraph
    node_types, edge_types = _types_for(small_ppi, 2)
    typed = CSRGraph(small_ppi.indptr, small_ppi.indices, node_types=node_types, edge_types=edge_types)
    kw = dict(embedding_size=8, epochs=1, walk_length=16, iterations=1, verbose=False, deterministic=True)
    model = Node2VecSkipGramB200(change_node_type_weight=0.05, change_edge_type_weight=2.0, **kw)
    assert model.is_using_node_types() and model.is_using_edge_types()
    a = model.fit_transform(typed, return_dataframe=False).get_all_node_embedding()[0]
    b = Node2VecSkipGramB200(**kw).fit_transform(typed, return_dataframe=False).get_all_node_embedding()[0]
    c = model.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()[0]  # untyped graph
    assert np.isfinite(a).all() and not np.array_equal(a, b) and np.array_equal(b, c)
