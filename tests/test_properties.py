"""Property tests of the CPU oracle on random graphs and kwargs (hypothesis; CPU only):
size-independent invariants of the path that hold for any input."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from embiggen_b200.graph import csr_from_edges, validate_csr

SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@st.composite
def graphs(draw):
    n = draw(st.integers(2, 40))
    m = draw(st.integers(1, 120))
    src = np.array(draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m)))
    dst = np.array(draw(st.lists(st.integers(0, n - 1), min_size=m, max_size=m)))
    directed = draw(st.booleans())
    graph = csr_from_edges(src, dst, n, symmetrise=not directed)
    return graph


weights = st.sampled_from([0.25, 0.5, 1.0, 2.0, 4.0, 7.5])


@settings(**SETTINGS)
@given(graphs(), weights, weights, st.integers(0, 2 ** 63), st.integers(2, 20), st.integers(0, 2 ** 40))
def test_walks_are_paths_of_the_graph(graph, rw, ew, seed, length, first):
    validate_csr(graph.indptr, graph.indices)
    if graph.indices.shape[0] == 0:
        return
    n = graph.get_number_of_nodes()
    srcs = oracle.sources(graph.indptr)
    walks, counters = oracle.walks(graph.indptr, graph.indices, seed, first, 50, length, rw, ew)
    edges = set(zip(np.repeat(np.arange(n), np.diff(graph.indptr)).tolist(), graph.indices.tolist()))
    for k, row in enumerate(walks):
        assert row[0] == srcs[(first + k) % len(srcs)]
        alive = True
        for a, b in zip(row[:-1], row[1:]):
            if not alive:
                assert b == oracle.PAD_TOKEN
            elif b == oracle.PAD_TOKEN:
                assert graph.indptr[a + 1] == graph.indptr[a]  # only a dead end stops a walk
                alive = False
            else:
                assert (int(a), int(b)) in edges
    assert counters["steps"] == int((walks[:, 1:] != oracle.PAD_TOKEN).sum())
    # a walk is a pure function of (seed, walk id): any other batching returns the same rows
    again, _ = oracle.walks(graph.indptr, graph.indices, seed, first + 7, 20, length, rw, ew)
    assert np.array_equal(again, walks[7:27])
    strided, _ = oracle.walks(graph.indptr, graph.indices, seed, first + 1, 16, length, rw, ew, walk_id_stride=3)
    assert np.array_equal(strided, walks[1:49:3])


@settings(**SETTINGS)
@given(graphs(), st.sampled_from([0.0, 0.5, 0.75, 1.0, 1.3]))
def test_alias_table_is_a_probability_table(graph, alpha):
    degrees = np.diff(graph.indptr).astype(np.float64)
    if degrees.sum() == 0:
        return
    thr, alias = oracle.alias_build(graph.indptr, alpha)
    n = len(degrees)
    keep = (thr.astype(np.float64) + (thr == 0xFFFFFFFF)) / 2.0 ** 32
    pmf = keep / n
    np.add.at(pmf, alias, (1.0 - keep) / n)
    target = np.where(degrees > 0, degrees ** alpha, 0.0)
    assert abs(pmf.sum() - 1.0) < 1e-9
    assert np.abs(pmf - target / target.sum()).max() < 1e-8
    assert (alias < n).all()


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(graphs(), st.sampled_from(["SkipGram", "CBOW"]), st.integers(1, 12), st.integers(0, 6), st.integers(1, 5),
       st.integers(2, 12), st.integers(0, 2 ** 62))
def test_training_is_deterministic_finite_and_counts_its_pairs(graph, model, D, K, w, L, seed):
    if graph.indices.shape[0] == 0:
        return
    n = graph.get_number_of_nodes()
    walks, _ = oracle.walks(graph.indptr, graph.indices, seed, 0, 12, L, 0.5, 2.0)
    thr, alias = oracle.alias_build(graph.indptr, 0.75)
    runs = []
    for _ in range(2):
        t0, t1 = oracle.init_tables(n, D, seed)
        stats = oracle.train(model, walks, t0, t1, seed, n, D, w, K, 0.05, thr=thr, alias=alias)
        runs.append((t0, t1, stats))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert runs[0][2] == runs[1][2]
    t0, t1, stats = runs[0]
    assert np.isfinite(t0).all() and np.isfinite(t1).all()
    # pairs = window positions holding a real token different from the centre
    expected = 0
    for row in walks:
        for i, c in enumerate(row):
            if c == oracle.PAD_TOKEN:
                break
            lo, hi = max(0, i - w), min(L - 1, i + w)
            expected += sum(1 for j in range(lo, hi + 1)
                            if j != i and row[j] != oracle.PAD_TOKEN and row[j] != c)
    assert stats["pairs"] == expected
    assert stats["targets"] <= (stats["pairs"] if model == "SkipGram" else 12 * L) * (K + 1)
    assert (t0[:, D:] == 0).all() and (t1[:, D:] == 0).all()


# ---- the widened path: alias tables, Walklets split, co-occurrence, typed / normalised walks ----
@settings(**SETTINGS)
@given(graphs(), st.integers(0, 2 ** 32 - 1), st.floats(0.0, 0.5))
def test_edge_alias_tables_encode_the_weights(graph, seed, zero_fraction):
    nnz = graph.indices.shape[0]
    if nnz == 0:
        return
    rng = np.random.default_rng(seed)
    weights = np.exp(rng.uniform(-8, 8, nnz)).astype(np.float32)
    weights[rng.random(nnz) < zero_fraction] = 0.0
    table = oracle.edge_alias(graph.indptr, weights)
    for v in range(graph.get_number_of_nodes()):
        lo, hi = graph.indptr[v], graph.indptr[v + 1]
        d = hi - lo
        if d == 0:
            continue
        thr, alias = table[lo:hi, 0].astype(np.float64), table[lo:hi, 1].astype(np.int64)
        assert (alias >= 0).all() and (alias < d).all()
        keep = np.where(thr >= 2.0 ** 32 - 1, 1.0, thr / 2.0 ** 32)
        pmf = keep / d
        np.add.at(pmf, alias, (1.0 - keep) / d)
        row = weights[lo:hi].astype(np.float64)
        expected = row / row.sum() if row.sum() > 0 else np.full(d, 1.0 / d)
        assert np.allclose(pmf, expected, atol=1e-9)
        if row.sum() > 0:
            assert (pmf[row == 0] == 0).all()


@settings(**SETTINGS)
@given(graphs(), st.integers(0, 2 ** 63), st.integers(2, 24), st.integers(1, 6), st.integers(1, 5))
def test_walklet_scales_partition_the_window(graph, seed, length, window, scale):
    """The pairs of the scales 1..w are disjoint and their union is the window-w pair set; the
    co-occurrence counts are symmetric and sum to the number of ordered pairs."""
    if graph.indices.shape[0] == 0:
        return
    walks, _ = oracle.walks(graph.indptr, graph.indices, seed, 0, 20, length)
    L = walks.shape[1]
    total = 0
    for k in range(1, min(window, L - 1) + 1):
        sub = oracle.walklet_split(walks, k)
        a, b = sub[:, :, :-1], sub[:, :, 1:]
        total += 2 * int(((a != oracle.PAD_TOKEN) & (b != oracle.PAD_TOKEN) & (a != b)).sum())
    centre, context, count = oracle.cooccurrence(walks, window)
    assert int(count.sum()) == total
    forward = {(int(c), int(o)): int(x) for c, o, x in zip(centre, context, count)}
    assert all(forward.get((o, c)) == x for (c, o), x in forward.items())
    if scale < L:
        sub = oracle.walklet_split(walks, scale)
        rebuilt = np.full((walks.shape[0], sub.shape[2] * scale), oracle.PAD_TOKEN, dtype=np.uint32)
        for r in range(scale):
            rebuilt[:, r::scale] = sub[r]
        assert np.array_equal(rebuilt[:, :L], walks)


@settings(**SETTINGS)
@given(graphs(), weights, weights, st.sampled_from([0.2, 1.0, 3.0]), st.sampled_from([0.5, 1.0, 4.0]),
       st.booleans(), st.booleans(), st.integers(0, 2 ** 63))
def test_typed_normalised_weighted_walks_are_paths(graph, rw, ew, cn, ce, weighted, normalize, seed):
    nnz = graph.indices.shape[0]
    if nnz == 0:
        return
    n = graph.get_number_of_nodes()
    rng = np.random.default_rng(seed % (2 ** 32))
    w = (rng.random(nnz) + 0.1).astype(np.float32) if weighted else None
    walks, counters = oracle.walks(graph.indptr, graph.indices, seed, 3, 40, 12, rw, ew, weights=w,
                                   normalize_by_degree=normalize, node_types=rng.integers(0, 3, n),
                                   edge_types=rng.integers(0, 3, nnz), change_node_type_weight=cn,
                                   change_edge_type_weight=ce)
    edges = set(zip(np.repeat(np.arange(n), np.diff(graph.indptr)).tolist(), graph.indices.tolist()))
    assert counters["capped"] == 0
    for row in walks:
        for a, b in zip(row[:-1], row[1:]):
            if b == oracle.PAD_TOKEN:
                break
            assert (int(a), int(b)) in edges
    again, _ = oracle.walks(graph.indptr, graph.indices, seed, 3, 40, 12, rw, ew, weights=w,
                            normalize_by_degree=normalize, node_types=rng.integers(0, 3, n),
                            edge_types=rng.integers(0, 3, nnz), change_node_type_weight=1.0,
                            change_edge_type_weight=1.0)
    assert again.shape == walks.shape  # unit change weights: the plain sampler, still valid walks
