"""Pins the CPU oracle's walk sampler (CPU only, no GPU).

The reference holds no golden walks for this path, so the oracle is pinned
against what the domain fixes: every transition is an edge, DeepWalk picks neighbours
uniformly, and second-order transitions follow the analytic node2vec distribution
(Grover & Leskovec 2016, eq. 2, with return_weight = 1/p and explore_weight = 1/q as in
/root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:58-71).
"""
import numpy as np
import pytest
from scipy import stats

import oracle
from conftest import tiny_graphs
from embiggen_b200.graph import csr_from_edges


def neighbours(graph, v):
    return graph.indices[graph.indptr[v]:graph.indptr[v + 1]]


def analytic_pmf(graph, prev, cur, rw, ew):
    """node2vec transition pmf out of `cur` having arrived from `prev`."""
    nv = neighbours(graph, cur)
    np_ = set(int(x) for x in neighbours(graph, prev))
    w = np.array([rw if x == prev else (1.0 if int(x) in np_ else ew) for x in nv], dtype=np.float64)
    return nv, w / w.sum()


def dense_test_graph():
    """12 nodes: a clique of 5, a ring, chords and a pendant (all three classes occur)."""
    src = [0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 4, 5, 6, 7, 8, 9, 10, 5, 6, 2, 11]
    dst = [1, 2, 3, 4, 2, 3, 4, 3, 4, 4, 5, 6, 7, 8, 9, 10, 4, 8, 10, 7, 0]
    return csr_from_edges(np.array(src), np.array(dst), 12, name="dense12")


@pytest.mark.parametrize("fixture", ["small_ppi", "er_graph", "rmat_graph"])
@pytest.mark.parametrize("rw,ew", [(1.0, 1.0), (0.25, 4.0), (2.0, 0.5)])
def test_walks_follow_edges_and_start_at_sources(request, fixture, rw, ew):
    graph = request.getfixturevalue(fixture)
    srcs = oracle.sources(graph.indptr)
    assert np.array_equal(srcs, np.flatnonzero(np.diff(graph.indptr) > 0).astype(np.uint32))
    first, count, L = 3, 2 * len(srcs) + 5, 40
    walks, counters = oracle.walks(graph.indptr, graph.indices, 11, first, count, L, rw, ew)
    assert walks.shape == (count, L) and walks.dtype == np.uint32
    assert np.array_equal(walks[:, 0], srcs[(first + np.arange(count)) % len(srcs)])
    n = graph.get_number_of_nodes()
    edge_keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(graph.indptr)) * n + graph.indices
    keys = walks[:, :-1].astype(np.int64).ravel() * n + walks[:, 1:].astype(np.int64).ravel()
    assert np.isin(keys, edge_keys).all()
    assert counters["steps"] == count * (L - 1) and counters["capped"] == 0
    if rw == 1.0 and ew == 1.0:
        assert counters["trials"] == 0 and counters["first_order"] == counters["steps"]
    else:
        assert counters["first_order"] == count  # only the first transition of every walk
        assert counters["trials"] >= counters["steps"] - count
        assert counters["searches"] <= counters["trials"]


def test_deepwalk_is_uniform_over_neighbours():
    graph = dense_test_graph()
    walks, _ = oracle.walks(graph.indptr, graph.indices, 5, 0, 120_000, 2)
    for v in (0, 4, 2):
        nxt = walks[walks[:, 0] == v, 1]
        nv = neighbours(graph, v)
        counts = np.array([(nxt == x).sum() for x in nv])
        assert counts.sum() == len(nxt) and len(nxt) > 5000
        assert stats.chisquare(counts).pvalue > 1e-3


@pytest.mark.parametrize("rw,ew", [(0.25, 4.0), (2.0, 0.5), (0.5, 2.0), (1.0, 3.0), (7.5, 1.0)])
def test_second_order_transitions_match_analytic_pmf(rw, ew):
    """chi-square of the third token given the first two against the analytic pmf."""
    graph = dense_test_graph()
    walks, _ = oracle.walks(graph.indptr, graph.indices, 2024, 0, 360_000, 3, rw, ew)
    checked = 0
    for prev, cur in [(0, 4), (4, 0), (1, 2), (5, 6), (11, 0), (2, 7), (10, 4)]:
        sel = walks[(walks[:, 0] == prev) & (walks[:, 1] == cur), 2]
        nv, pmf = analytic_pmf(graph, prev, cur, rw, ew)
        counts = np.array([(sel == x).sum() for x in nv])
        assert counts.sum() == len(sel)
        if len(sel) < 500:
            continue
        assert stats.chisquare(counts, pmf * len(sel)).pvalue > 1e-3, (prev, cur)
        checked += 1
    assert checked >= 5


def test_second_order_on_small_ppi_hub(small_ppi):
    """The hub of BASELINE config C1's graph (degree 348): pooled chi-square by class."""
    rw, ew = 0.25, 4.0
    degrees = np.diff(small_ppi.indptr)
    hub = int(degrees.argmax())
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 1, 0, 300_000, 3, rw, ew)
    sel = walks[walks[:, 1] == hub]
    observed = np.zeros(3)
    expected = np.zeros(3)
    for prev in np.unique(sel[:, 0]):
        rows = sel[sel[:, 0] == prev]
        nv, pmf = analytic_pmf(small_ppi, int(prev), hub, rw, ew)
        np_ = set(int(x) for x in neighbours(small_ppi, int(prev)))
        cls = np.array([0 if x == prev else (1 if int(x) in np_ else 2) for x in nv])
        for c in range(3):
            expected[c] += pmf[cls == c].sum() * len(rows)
            observed[c] += np.isin(rows[:, 2], nv[cls == c]).sum()
    keep = expected > 5
    assert keep.sum() >= 2 and observed.sum() > 10_000
    assert stats.chisquare(observed[keep], expected[keep] * observed[keep].sum() / expected[keep].sum()).pvalue > 1e-3


def test_thresholds_are_exact_integer_ratios():
    thr = oracle.thresholds(0.25, 4.0)
    assert list(thr) == [2 ** 32 // 16, 2 ** 32 // 4, 2 ** 32]
    thr = oracle.thresholds(2.0, 0.5)
    assert list(thr) == [2 ** 32, 2 ** 31, 2 ** 30]
    assert list(oracle.thresholds(1.0, 1.0)) == [2 ** 32] * 3


def test_walk_ids_are_independent_of_batching(er_graph):
    whole, _ = oracle.walks(er_graph.indptr, er_graph.indices, 42, 0, 3000, 24, 0.5, 2.0)
    for world in (2, 3, 8):
        for rank in range(world):
            count = (3000 - rank + world - 1) // world
            shard, _ = oracle.walks(er_graph.indptr, er_graph.indices, 42, rank, count, 24, 0.5,
                                    2.0, walk_id_stride=world)
            assert np.array_equal(shard, whole[rank::world])
    tail, _ = oracle.walks(er_graph.indptr, er_graph.indices, 42, 2500, 500, 24, 0.5, 2.0)
    assert np.array_equal(tail, whole[2500:])
    other, _ = oracle.walks(er_graph.indptr, er_graph.indices, 43, 0, 3000, 24, 0.5, 2.0)
    assert not np.array_equal(other, whole)


def test_multithreaded_oracle_gives_the_same_walks(er_graph):
    single, c1 = oracle.walks(er_graph.indptr, er_graph.indices, 7, 0, 5000, 32, 0.25, 4.0)
    oracle.set_threads(4)
    try:
        multi, c4 = oracle.walks(er_graph.indptr, er_graph.indices, 7, 0, 5000, 32, 0.25, 4.0)
    finally:
        oracle.set_threads(1)
    assert np.array_equal(single, multi) and c1 == c4


@pytest.mark.parametrize("name", sorted(tiny_graphs()))
def test_edge_case_graphs(name):
    graph = tiny_graphs()[name]
    for rw, ew in [(1.0, 1.0), (0.25, 4.0), (4.0, 0.25)]:
        walks, counters = oracle.walks(graph.indptr, graph.indices, 3, 0, 64, 16, rw, ew)
        alive = walks != oracle.PAD_TOKEN
        assert alive[:, 0].all()
        # PAD only ever follows PAD or a dead end
        for row, mask in zip(walks, alive):
            stop = int(mask.sum())
            assert mask[:stop].all() and not mask[stop:].any()
            if stop < 16:
                last = int(row[stop - 1])
                assert graph.indptr[last + 1] == graph.indptr[last]
        if name == "directed_dead_end":
            assert (~alive).any()
        else:
            assert alive.all()
        if name == "path" and (rw, ew) == (1.0, 1.0):
            assert (np.abs(np.diff(walks.astype(np.int64), axis=1)) == 1).all()
        if name == "star":
            hub_positions = walks[walks[:, 0] == 0][:, ::2]
            assert (hub_positions == 0).all()


# ---- edge weights: proposals proportional to the weights, p/q bias on top (KnightKing) ----
def weighted_test_graph():
    graph = dense_test_graph()
    rng = np.random.default_rng(4)
    n = graph.get_number_of_nodes()
    rows = np.repeat(np.arange(n), np.diff(graph.indptr))
    # symmetric weights spanning three orders of magnitude, one zero-weight edge (never taken)
    table = np.exp(rng.uniform(-3, 3, size=(n, n)))
    table = np.minimum(table, table.T)
    weights = table[rows, graph.indices].astype(np.float32)
    weights[(rows == 0) & (graph.indices == 11)] = 0.0
    weights[(rows == 11) & (graph.indices == 0)] = 0.0
    return graph, weights


def test_edge_alias_tables_reproduce_the_weights():
    """pmf implied by a row's alias table == weights / total (to 2^-31 per entry); zero-weight
    edges are never proposed; a row of total weight zero is uniform."""
    graph, weights = weighted_test_graph()
    table = oracle.edge_alias(graph.indptr, weights)
    assert table.shape == (graph.indices.shape[0], 2) and table.dtype == np.uint32
    for v in range(graph.get_number_of_nodes()):
        lo, hi = graph.indptr[v], graph.indptr[v + 1]
        d = hi - lo
        thr, alias = table[lo:hi, 0].astype(np.float64), table[lo:hi, 1].astype(np.int64)
        assert (alias < d).all()
        # coin: low word < thr keeps the slot (thr = 2^32 - 1 means "always" up to 2^-32)
        keep = np.where(thr >= 2.0 ** 32 - 1, 1.0, thr / 2.0 ** 32)
        pmf = keep / d
        np.add.at(pmf, alias, (1.0 - keep) / d)
        row = weights[lo:hi].astype(np.float64)
        if row.sum() == 0:                     # node 11: its only edge weighs zero -> uniform row
            assert np.allclose(pmf, 1.0 / d)
            continue
        assert np.allclose(pmf, row / row.sum(), atol=1e-9)
        assert (pmf[row == 0] == 0).all()
    with pytest.raises(ValueError):
        oracle.edge_alias(graph.indptr, -weights)
    zero = oracle.edge_alias(graph.indptr, np.zeros_like(weights))
    assert (zero[:, 0] == 0xFFFFFFFF).all()   # uniform rows


def test_weighted_first_order_follows_the_weights():
    graph, weights = weighted_test_graph()
    walks, _ = oracle.walks(graph.indptr, graph.indices, 8, 0, 240_000, 2, weights=weights)
    for v in (0, 4, 2, 7):
        lo, hi = graph.indptr[v], graph.indptr[v + 1]
        nxt = walks[walks[:, 0] == v, 1]
        counts = np.array([(nxt == x).sum() for x in graph.indices[lo:hi]])
        pmf = weights[lo:hi].astype(np.float64) / weights[lo:hi].sum()
        assert counts[pmf == 0].sum() == 0
        keep = pmf * len(nxt) >= 5
        assert keep.sum() >= 2
        expected = pmf[keep] * len(nxt)
        assert stats.chisquare(counts[keep], expected * counts[keep].sum() / expected.sum()).pvalue > 1e-3


@pytest.mark.parametrize("rw,ew", [(0.25, 4.0), (2.0, 0.5)])
def test_weighted_second_order_matches_analytic_pmf(rw, ew):
    graph, weights = weighted_test_graph()
    walks, _ = oracle.walks(graph.indptr, graph.indices, 77, 0, 480_000, 3, rw, ew, weights=weights)
    checked = 0
    for prev, cur in [(4, 0), (0, 4), (1, 2), (5, 6), (2, 7)]:
        sel = walks[(walks[:, 0] == prev) & (walks[:, 1] == cur), 2]
        lo, hi = graph.indptr[cur], graph.indptr[cur + 1]
        nv, bias = analytic_pmf(graph, prev, cur, rw, ew)  # pure p/q bias ...
        pmf = bias * weights[lo:hi]                         # ... times the static weight
        pmf = pmf / pmf.sum()
        counts = np.array([(sel == x).sum() for x in nv])
        keep = pmf * len(sel) >= 5
        if len(sel) < 500 or keep.sum() < 2:
            continue
        expected = pmf[keep] * len(sel)
        assert stats.chisquare(counts[keep], expected * counts[keep].sum() / expected.sum()).pvalue > 1e-3
        checked += 1
    assert checked >= 3


def test_unit_weights_reproduce_nothing_else_than_valid_walks(small_ppi_weighted):
    g = small_ppi_weighted
    walks, counters = oracle.walks(g.indptr, g.indices, 5, 0, 2000, 24, 0.25, 4.0, weights=g.weights)
    n = g.get_number_of_nodes()
    edge_keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(g.indptr)) * n + g.indices
    keys = walks[:, :-1].astype(np.int64).ravel() * n + walks[:, 1:].astype(np.int64).ravel()
    assert np.isin(keys, edge_keys).all() and counters["steps"] == 2000 * 23
    unweighted, _ = oracle.walks(g.indptr, g.indices, 5, 0, 2000, 24, 0.25, 4.0)
    assert not np.array_equal(walks, unweighted)


# ---- normalize_by_degree: transition weight / deg(destination) (node2vec_skipgram.py:94-96) ----
def test_degree_normalised_weights():
    graph = dense_test_graph()
    degrees = np.diff(graph.indptr).astype(np.float32)
    got = oracle.degree_normalised_weights(graph.indptr, graph.indices)
    assert got.dtype == np.float32 and np.array_equal(got, np.float32(1.0) / degrees[graph.indices])
    dead = tiny_graphs()["directed_dead_end"]           # a dead end weighs like a leaf
    got = oracle.degree_normalised_weights(dead.indptr, dead.indices, np.full(dead.indices.shape[0], 3.0))
    assert np.array_equal(got, np.float32(3.0) / np.maximum(np.diff(dead.indptr), 1).astype(np.float32)[dead.indices])


@pytest.mark.parametrize("rw,ew", [(1.0, 1.0), (0.25, 4.0), (2.0, 0.5)])
@pytest.mark.parametrize("weighted", [False, True])
def test_normalize_by_degree_matches_analytic_pmf(rw, ew, weighted):
    """First transition ~ w(v,x) / deg(x); later ones ~ bias(prev,x) * w(v,x) / deg(x)."""
    graph, weights = weighted_test_graph()
    if not weighted:
        weights = None
    degrees = np.diff(graph.indptr).astype(np.float64)
    walks, counters = oracle.walks(graph.indptr, graph.indices, 31, 0, 480_000, 3, rw, ew,
                                   weights=weights, normalize_by_degree=True)
    assert counters["capped"] == 0
    if (rw, ew) == (1.0, 1.0):  # folded into the proposal: no trial at all for a first-order walk
        assert counters["trials"] == 0
    for v in (0, 4, 2):
        lo, hi = graph.indptr[v], graph.indptr[v + 1]
        nxt = walks[walks[:, 0] == v, 1]
        pmf = (1.0 if weights is None else weights[lo:hi].astype(np.float64)) / degrees[graph.indices[lo:hi]]
        pmf = pmf / pmf.sum()
        counts = np.array([(nxt == x).sum() for x in graph.indices[lo:hi]])
        keep = pmf * len(nxt) >= 5
        expected = pmf[keep] * len(nxt)
        assert stats.chisquare(counts[keep], expected * counts[keep].sum() / expected.sum()).pvalue > 1e-3
    checked = 0
    for prev, cur in [(4, 0), (0, 4), (1, 2), (5, 6), (2, 7)]:
        sel = walks[(walks[:, 0] == prev) & (walks[:, 1] == cur), 2]
        lo, hi = graph.indptr[cur], graph.indptr[cur + 1]
        nv, bias = analytic_pmf(graph, prev, cur, rw, ew)
        pmf = bias * (1.0 if weights is None else weights[lo:hi]) / degrees[nv]
        pmf = pmf / pmf.sum()
        counts = np.array([(sel == x).sum() for x in nv])
        keep = pmf * len(sel) >= 5
        if len(sel) < 500 or keep.sum() < 2:
            continue
        expected = pmf[keep] * len(sel)
        assert stats.chisquare(counts[keep], expected * counts[keep].sum() / expected.sum()).pvalue > 1e-3
        checked += 1
    assert checked >= 3


@pytest.mark.parametrize("name", sorted(tiny_graphs()))
def test_normalize_by_degree_edge_case_graphs(name):
    graph = tiny_graphs()[name]
    if (np.diff(graph.indptr) > 0).sum() == 0:
        return
    walks, counters = oracle.walks(graph.indptr, graph.indices, 3, 0, 200, 9, 0.25, 4.0,
                                   normalize_by_degree=True)
    n = graph.get_number_of_nodes()
    edge_keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(graph.indptr)) * n + graph.indices
    a, b = walks[:, :-1].ravel(), walks[:, 1:].ravel()
    live = b != 0xFFFFFFFF
    assert np.isin(a[live].astype(np.int64) * n + b[live], edge_keys).all()
    assert counters["capped"] == 0


# ---- typed walks: change_node_type_weight / change_edge_type_weight (node2vec_skipgram.py:72-77) ----
def typed_test_graph():
    graph, weights = weighted_test_graph()
    rng = np.random.default_rng(9)
    n = graph.get_number_of_nodes()
    node_types = rng.integers(0, 3, n).astype(np.uint32)
    rows = np.repeat(np.arange(n), np.diff(graph.indptr))
    table = rng.integers(0, 3, (n, n))
    table = np.triu(table) + np.triu(table, 1).T           # an undirected edge has one type
    edge_types = table[rows, graph.indices].astype(np.uint32)
    return graph, weights, node_types, edge_types


def typed_pmf(graph, weights, node_types, edge_types, prev, cur, rw, ew, cn, ce, normalize):
    lo, hi = graph.indptr[cur], graph.indptr[cur + 1]
    nv = graph.indices[lo:hi]
    w = np.ones(len(nv)) if weights is None else weights[lo:hi].astype(np.float64)
    if prev is not None:
        _, bias = analytic_pmf(graph, prev, cur, rw, ew)
        plo = graph.indptr[prev]
        back = plo + int(np.searchsorted(neighbours(graph, prev), cur))
        w = w * bias * np.where(edge_types[lo:hi] != edge_types[back], ce, 1.0)
    w = w * np.where(node_types[nv] != node_types[cur], cn, 1.0)
    if normalize:
        w = w / np.diff(graph.indptr)[nv]
    return nv, w / w.sum()


@pytest.mark.parametrize("rw,ew,cn,ce,weighted,normalize", [
    (1.0, 1.0, 3.0, 1.0, False, False), (1.0, 1.0, 1.0, 0.25, False, False),
    (0.25, 4.0, 0.2, 5.0, False, False), (2.0, 0.5, 4.0, 0.5, True, False),
    (0.5, 2.0, 0.5, 2.0, True, True)])
def test_typed_walks_match_analytic_pmf(rw, ew, cn, ce, weighted, normalize):
    graph, weights, node_types, edge_types = typed_test_graph()
    if not weighted:
        weights = None
    walks, counters = oracle.walks(graph.indptr, graph.indices, 5, 0, 480_000, 3, rw, ew, weights=weights,
                                   normalize_by_degree=normalize, node_types=node_types,
                                   edge_types=edge_types, change_node_type_weight=cn,
                                   change_edge_type_weight=ce)
    assert counters["capped"] == 0

    def check(sel, pmf, nv):
        counts = np.array([(sel == x).sum() for x in nv])
        keep = pmf * len(sel) >= 5
        if len(sel) < 500 or keep.sum() < 2:
            return 0
        expected = pmf[keep] * len(sel)
        assert stats.chisquare(counts[keep], expected * counts[keep].sum() / expected.sum()).pvalue > 1e-3
        return 1

    checked = 0
    for v in (0, 4, 2):
        nv, pmf = typed_pmf(graph, weights, node_types, edge_types, None, v, rw, ew, cn, ce, normalize)
        checked += check(walks[walks[:, 0] == v, 1], pmf, nv)
    for prev, cur in [(4, 0), (0, 4), (1, 2), (5, 6), (2, 7)]:
        nv, pmf = typed_pmf(graph, weights, node_types, edge_types, prev, cur, rw, ew, cn, ce, normalize)
        checked += check(walks[(walks[:, 0] == prev) & (walks[:, 1] == cur), 2], pmf, nv)
    assert checked >= 6


def test_unit_type_weights_leave_the_walks_unchanged(small_ppi):
    n = small_ppi.get_number_of_nodes()
    node_types = (np.arange(n) % 4).astype(np.uint32)
    edge_types = (np.arange(small_ppi.indices.shape[0]) % 3).astype(np.uint32)
    plain, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 1, 0, 500, 20, 0.25, 4.0)
    typed, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 1, 0, 500, 20, 0.25, 4.0,
                            node_types=node_types, edge_types=edge_types)
    assert np.array_equal(plain, typed)
    changed, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 1, 0, 500, 20, 0.25, 4.0,
                              node_types=node_types, change_node_type_weight=0.1)
    assert not np.array_equal(plain, changed)
    # a very small change weight keeps the walk inside the type of its start node when it can
    stay = (node_types[changed[:, 1:]] == node_types[changed[:, :-1]]).mean()
    assert stay > (node_types[plain[:, 1:]] == node_types[plain[:, :-1]]).mean() + 0.2


# ---- folded return edge (oracle/walks.c: orc_fold_thresholds) ----
def test_fold_thresholds_and_symmetry_probe(small_ppi, rmat_graph):
    thrf, excess = oracle.fold_thresholds(2.0, 0.5)  # C3: p = 0.5, q = 2
    assert list(thrf) == [2 ** 32, 2 ** 32, 2 ** 31] and excess == 2 ** 20
    thrf, excess = oracle.fold_thresholds(7.5, 1.0)
    assert list(thrf) == [2 ** 32] * 3 and excess == int(6.5 * 2 ** 20)
    thrf, excess = oracle.fold_thresholds(6.0, 4.0)  # envelope 4: common 1/4, explore 1, excess 1/2
    assert list(thrf) == [2 ** 32, 2 ** 30, 2 ** 32] and excess == 2 ** 19
    for rw, ew in [(0.25, 4.0), (1.0, 1.0), (0.5, 2.0), (1.0, 3.0), (3.0, 3.0)]:
        assert oracle.fold_thresholds(rw, ew)[1] == 0  # nothing to fold: the plain envelope stays
    assert oracle.is_undirected(small_ppi.indptr, small_ppi.indices)
    assert oracle.is_undirected(rmat_graph.indptr, rmat_graph.indices)
    directed = tiny_graphs()["directed_dead_end"]
    assert not oracle.is_undirected(directed.indptr, directed.indices)


@pytest.mark.parametrize("rw,ew", [(2.0, 0.5), (7.5, 1.0), (6.0, 4.0)])
def test_folded_sampler_follows_the_same_pmf_in_fewer_trials(rmat_graph, rw, ew):
    """Same analytic pmf (test_second_order_transitions_match_analytic_pmf covers the dense graph
    with the fold on, since it is undirected); here: the plain and the folded streams agree in
    distribution on a skewed graph and the fold needs fewer proposals."""
    g = rmat_graph
    folded, fc = oracle.walks(g.indptr, g.indices, 5, 0, 60_000, 3, rw, ew, undirected=True)
    plain, pc = oracle.walks(g.indptr, g.indices, 5, 0, 60_000, 3, rw, ew, undirected=False)
    assert fc["steps"] == pc["steps"] and fc["trials"] < pc["trials"]
    # class frequencies of the third token (return / other) agree between the two samplers
    back = [(w[:, 2] == w[:, 0]).sum() for w in (folded, plain)]
    table = np.array([[back[0], len(folded) - back[0]], [back[1], len(plain) - back[1]]])
    assert stats.chi2_contingency(table)[1] > 1e-3
    # a directed graph never folds, whatever the caller claims about p/q
    directed = tiny_graphs()["directed_dead_end"]
    a, _ = oracle.walks(directed.indptr, directed.indices, 5, 0, 200, 8, rw, ew)
    b, _ = oracle.walks(directed.indptr, directed.indices, 5, 0, 200, 8, rw, ew, undirected=False)
    assert np.array_equal(a, b)


def test_synthetic_graph_generator_matches_the_numpy_definition():
    """oracle/graphgen.c (OpenMP, hash set) == embiggen_b200/graph.py (numpy, sort): the first m
    distinct undirected edges of the Philox stream, whatever the thread count."""
    from embiggen_b200.graph import erdos_renyi, rmat
    for threads in (1, 4):
        oracle.set_threads(threads)
        try:
            for n, m in [(1000, 5000), (50, 600), (3, 1)]:
                indptr, indices = oracle.synthetic_csr("er", n, m)
                g = erdos_renyi(n, m, seed=42)
                assert np.array_equal(indptr, g.indptr) and np.array_equal(indices, g.indices)
            for scale, n, m in [(10, 900, 4000), (12, 4096, 30000), (14, 10000, 100_000)]:
                indptr, indices = oracle.synthetic_csr("rmat", n, m, scale=scale, seed=11)
                g = rmat(scale, m, n=n, seed=11)
                assert np.array_equal(indptr, g.indptr) and np.array_equal(indices, g.indices)
        finally:
            oracle.set_threads(1)
    with pytest.raises(ValueError):
        oracle.synthetic_csr("er", 10, 40)
