"""GloVe on walk co-occurrences against the CPU oracle, through the C ABI:
co-occurrence triples integer-exact, the single-warp SGD launch bit-exact, the production
launch by tolerance."""
import numpy as np
import pytest

import oracle
from embiggen_b200.engine import Engine

pytestmark = pytest.mark.gpu
SEED = 42


@pytest.mark.parametrize("rw,ew,L,w", [(1.0, 1.0, 40, 5), (0.25, 4.0, 33, 3), (2.0, 0.5, 8, 10)])
def test_cooccurrence_exact(small_ppi, rmat_graph, rw, ew, L, w):
    for graph in (small_ppi, rmat_graph):
        n_walks = 2 * int((np.diff(graph.indptr) > 0).sum()) + 7
        walks, _ = oracle.walks(graph.indptr, graph.indices, SEED, 11, n_walks, L, rw, ew)
        expected = oracle.cooccurrence(walks, w)
        # chunk_walks forces several accumulate-and-merge rounds
        with Engine("GloVe", walk_length=L, window_size=w, return_weight=rw, explore_weight=ew,
                    iterations=1, chunk_walks=1000) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            assert engine.cooccurrence(SEED, 11, n_walks) == expected[0].shape[0]
            got = engine.export_cooccurrence()
            for a, b in zip(got, expected):
                assert np.array_equal(a, b)
            # two halves accumulated = the whole
            half = n_walks // 2
            engine.cooccurrence(SEED, 11, half)
            engine.cooccurrence(SEED, 11 + half, n_walks - half, accumulate=True)
            for a, b in zip(engine.export_cooccurrence(), expected):
                assert np.array_equal(a, b)


@pytest.mark.parametrize("D,alpha", [(100, 0.75), (8, 0.5), (200, 1.0), (300, 0.0)])
def test_glove_deterministic_bit_exact(small_ppi, D, alpha):
    n, L, w, lr, n_walks = small_ppi.get_number_of_nodes(), 32, 4, 0.05, 1064
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, SEED, 0, n_walks, L, 0.25, 4.0)
    centre, context, count = oracle.cooccurrence(walks, w)
    t0, t1 = oracle.init_tables(n, D, SEED)
    t0 *= 20.0  # leave the linear regime
    t1 *= 20.0
    with Engine("GloVe", embedding_size=D, walk_length=L, window_size=w, return_weight=0.25,
                explore_weight=4.0, iterations=1, glove_alpha=alpha, deterministic=True) as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        engine.import_tables(t0[:, :D], t1[:, :D])
        engine.cooccurrence(SEED, 0, n_walks)
        engine.reset_counters()
        for _ in range(2):
            engine.glove_train(lr)
        g0, g1 = engine.export_tables()
        counters = engine.counters()
    stats = {"loss_sum": 0.0, "trained": 0}
    for _ in range(2):
        r = oracle.glove_train(centre, context, count, t0, t1, D, alpha, lr)
        stats["loss_sum"] += r["loss_sum"]
        stats["trained"] += r["trained"]
    assert counters["pairs"] == stats["trained"] > 0
    assert np.array_equal(g0, t0[:, :D]) and np.array_equal(g1, t1[:, :D])
    assert np.isclose(counters["loss_sum"], stats["loss_sum"], rtol=1e-4)


def test_glove_fit_tracks_the_oracle(small_ppi):
    kw = dict(embedding_size=32, epochs=8, walk_length=64, window_size=4, learning_rate=0.05,
              learning_rate_decay=0.9)
    _, _, expected = oracle.glove_fit(small_ppi.indptr, small_ppi.indices, 5, alpha=0.75,
                                      return_weight=0.25, explore_weight=4.0, **kw)
    with Engine("GloVe", return_weight=0.25, explore_weight=4.0, iterations=1, **kw) as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        c, x, got = engine.fit(5)
    print("oracle", np.round(expected, 4), "gpu", np.round(got, 4))
    assert np.isfinite(c).all() and np.isfinite(x).all()
    assert got[-1] < 0.5 * got[0]
    # concurrent centres (atomic row updates, stale reads) against the sequential oracle:
    # stated tolerance: 20 % on the first three epochs (tiles of one centre's row train
    # concurrently from the same snapshot), 5 % afterwards (measured: 14 %, 11 %, then < 4 %)
    for epoch, (a, b) in enumerate(zip(expected, got)):
        assert abs(a - b) <= (0.20 if epoch < 3 else 0.05) * a, (epoch, a, b)
    # deterministic launch: the same tables as the oracle, whole path
    t0, t1, _ = oracle.glove_fit(small_ppi.indptr, small_ppi.indices, 5, alpha=0.75,
                                 return_weight=0.25, explore_weight=4.0, **dict(kw, epochs=2))
    with Engine("GloVe", return_weight=0.25, explore_weight=4.0, iterations=1, deterministic=True,
                **dict(kw, epochs=2)) as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        c, x, _ = engine.fit(5)
    assert np.array_equal(c, t0[:, :32]) and np.array_equal(x, t1[:, :32])


def test_glove_embedders(small_ppi):
    from embiggen_b200.embedders import DeepWalkGloVeB200, Node2VecGloVeB200
    for cls in (Node2VecGloVeB200, DeepWalkGloVeB200):
        model = cls(embedding_size=16, epochs=4, walk_length=32, verbose=False)
        tables = model.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
        assert len(tables) == 2 and all(t.shape == (1064, 16) and np.isfinite(t).all() for t in tables)
        losses = model.get_losses()
        assert len(losses) == 4 and losses[-1] < losses[0]


def test_glove_error_paths(small_ppi):
    with Engine("GloVe") as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        with pytest.raises(RuntimeError):
            engine.glove_train(0.05)  # no co-occurrence yet
        with pytest.raises(RuntimeError):
            engine.train_chunk(SEED, 0, 0.05)
    with Engine("SkipGram") as engine:
        engine.load_csr(small_ppi.indptr, small_ppi.indices)
        with pytest.raises(RuntimeError):
            engine.glove_train(0.05)
    with pytest.raises(ValueError):
        Engine("GloVe", walklet_scale=2)


def test_glove_by_centre_ranges_equals_the_one_piece_epoch(monkeypatch, small_ppi, rmat_graph):
    """Past 2^31 key slots an epoch's co-occurrence is counted and trained by ranges of centre ids
    (b2e_api.cu: glove_epoch_by_ranges).  B2E_GLOVE_SLOTS forces that path on small graphs: in the
    single-warp launch the tables equal the one-piece path and the oracle bit for bit, whatever
    the number of ranges (exact counts, x_max of the whole epoch, centres in ascending order)."""
    kw = dict(embedding_size=24, epochs=2, walk_length=40, window_size=3, learning_rate=0.05, learning_rate_decay=0.9)
    for graph in (small_ppi, rmat_graph):
        t0, t1, expected = oracle.glove_fit(graph.indptr, graph.indices, 9, alpha=0.75, return_weight=0.5,
                                            explore_weight=2.0, **kw)
        for slots in (None, 200_000, 3_000):
            if slots is None:
                monkeypatch.delenv("B2E_GLOVE_SLOTS", raising=False)
            else:
                monkeypatch.setenv("B2E_GLOVE_SLOTS", str(slots))
            with Engine("GloVe", return_weight=0.5, explore_weight=2.0, iterations=1, deterministic=True, **kw) as engine:
                engine.load_csr(graph.indptr, graph.indices)
                c, x, losses = engine.fit(9)
            assert np.array_equal(c, t0[:, :24]) and np.array_equal(x, t1[:, :24]), slots
            assert np.allclose(losses, expected, rtol=2e-3)  # float32 loss sums, another order
        # the production launch on many ranges: finite, and the loss falls like the one-piece run's
        monkeypatch.setenv("B2E_GLOVE_SLOTS", "50000")
        with Engine("GloVe", return_weight=0.5, explore_weight=2.0, iterations=1, **dict(kw, epochs=6)) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            c, x, ranged = engine.fit(9)
        monkeypatch.delenv("B2E_GLOVE_SLOTS")
        with Engine("GloVe", return_weight=0.5, explore_weight=2.0, iterations=1, **dict(kw, epochs=6)) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            _, _, whole = engine.fit(9)
        assert np.isfinite(c).all() and np.isfinite(x).all() and ranged[-1] < 0.6 * ranged[0]
        assert abs(ranged[-1] - whole[-1]) <= 0.1 * whole[-1]
