"""Downstream leg of the parity contract (north_star, SURVEY.md 8c leg 4): edge-prediction AUROC
of the GPU embedding against the embedding of the CPU oracle trained on the same graph with the
same kwargs.  The reference runs this comparison with `edge_prediction_evaluation`
(/root/reference/embiggen/edge_prediction/edge_prediction_evaluation.py:12-200: connected
Monte-Carlo holdout, embedder fitted on the training graph, sklearn-style classifier on
Hadamard edge features, `binary_auroc`); without the `ensmallen` wheel the same protocol is
restated here with numpy + scikit-learn and the oracle stands where Ensmallen would.

Tolerance (stated, as the contract asks): over three holdouts the mean AUROC of the GPU embedding
must be within 0.005 of the oracle's, two-sided (the 0.005 of north_star), and no single holdout
may differ by more than 0.01 either way (measured in round 2: every holdout within 0.002, means
+0.0005 ... +0.0011; the comparator is the 8-thread Hogwild oracle, itself not deterministic).
"""
import numpy as np
import pytest
from sklearn.linear_model import LogisticRegression
from sklearn.metrics import roc_auc_score

import oracle
from embiggen_b200.graph import csr_from_edges

KW = dict(embedding_size=32, walk_length=32, window_size=4, iterations=3, epochs=4,
          number_of_negative_samples=5, learning_rate=0.05, learning_rate_decay=0.9)


def block_model(seed, n=6000, block=100, degree_in=10, degree_out=2):
    """Planted-partition graph: blocks of `block` nodes with about `degree_in` neighbours inside
    and `degree_out` anywhere; returns the (deduplicated, upper-triangular) edge list."""
    rng = np.random.default_rng(seed)
    src, dst = [], []
    for b in range(n // block):
        m = block * degree_in // 2
        src.append(rng.integers(0, block, m) + b * block)
        dst.append(rng.integers(0, block, m) + b * block)
    m = n * degree_out // 2
    src.append(rng.integers(0, n, m))
    dst.append(rng.integers(0, n, m))
    src, dst = np.concatenate(src), np.concatenate(dst)
    keep = src != dst
    lo, hi = np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])
    keys = np.unique(lo * n + hi)
    return keys // n, keys % n, n


def holdout(src, dst, n, seed, train_fraction=0.8):
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(src))
    cut = int(train_fraction * len(src))
    train, test = order[:cut], order[cut:]
    existing = src * n + dst

    def negatives(count):
        a, b = rng.integers(0, n, 2 * count), rng.integers(0, n, 2 * count)
        ok = (a != b) & ~np.isin(np.minimum(a, b) * n + np.maximum(a, b), existing)
        return np.stack([a[ok][:count], b[ok][:count]], axis=1)

    return (src[train], dst[train]), (src[test], dst[test]), negatives(len(train)), negatives(len(test))


def auroc(embedding, train_pos, test_pos, train_neg, test_neg):
    """Hadamard edge features + logistic regression, AUROC on the held-out edges."""
    def features(a, b):
        return embedding[a] * embedding[b]
    x_train = np.vstack([features(*train_pos), features(train_neg[:, 0], train_neg[:, 1])])
    y_train = np.r_[np.ones(len(train_pos[0])), np.zeros(len(train_neg))]
    x_test = np.vstack([features(*test_pos), features(test_neg[:, 0], test_neg[:, 1])])
    y_test = np.r_[np.ones(len(test_pos[0])), np.zeros(len(test_neg))]
    scale = np.abs(x_train).max() + 1e-12
    model = LogisticRegression(max_iter=500, C=10.0).fit(x_train / scale, y_train)
    return roc_auc_score(y_test, model.decision_function(x_test / scale))


def oracle_embedding(model, graph, seed, rw, ew, threads, shared_negatives=False):
    oracle.set_threads(threads)
    try:
        t0, t1, _ = oracle.fit(model, graph.indptr, graph.indices, seed, KW["embedding_size"], KW["epochs"],
                               KW["iterations"], KW["walk_length"], KW["window_size"],
                               KW["number_of_negative_samples"], KW["learning_rate"],
                               KW["learning_rate_decay"], return_weight=rw, explore_weight=ew,
                               shared_negatives=shared_negatives)
    finally:
        oracle.set_threads(1)
    D = KW["embedding_size"]
    return np.hstack([t0[:, :D], t1[:, :D]])  # classifiers hstack both tables (node_transformer.py:108-116)


def test_oracle_embedding_predicts_held_out_edges():
    """CPU only: the normative algorithm learns the planted structure (sanity of the spec)."""
    src, dst, n = block_model(0)
    train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, 0)
    graph = csr_from_edges(train_pos[0], train_pos[1], n)
    embedding = oracle_embedding("SkipGram", graph, 42, 1.0, 1.0, threads=8)
    assert auroc(embedding, train_pos, test_pos, train_neg, test_neg) > 0.85


def test_oracle_shared_negatives_embedding_predicts_held_out_edges():
    """CPU only: the opt-in estimator with one set of negatives per centre (oracle/sgns.c:
    train_centre_shared) learns the same planted structure, within 0.02 of the per-pair oracle."""
    src, dst, n = block_model(0)
    train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, 0)
    graph = csr_from_edges(train_pos[0], train_pos[1], n)
    plain = auroc(oracle_embedding("SkipGram", graph, 42, 1.0, 1.0, threads=8), train_pos, test_pos, train_neg, test_neg)
    shared = auroc(oracle_embedding("SkipGram", graph, 42, 1.0, 1.0, threads=8, shared_negatives=True),
                   train_pos, test_pos, train_neg, test_neg)
    print("oracle AUROC: per-pair negatives", plain, "shared negatives", shared)
    assert shared > 0.85 and shared > plain - 0.02


@pytest.mark.gpu
@pytest.mark.parametrize("model,rw,ew", [("SkipGram", 1.0, 1.0), ("SkipGram", 0.25, 4.0), ("CBOW", 2.0, 0.5)])
def test_gpu_auroc_matches_oracle_auroc(model, rw, ew):
    from embiggen_b200.engine import Engine
    deltas = []
    for trial in range(3):
        src, dst, n = block_model(trial)
        train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, trial)
        graph = csr_from_edges(train_pos[0], train_pos[1], n)
        reference = oracle_embedding(model, graph, 42 + trial, rw, ew, threads=8)
        with Engine(model, return_weight=rw, explore_weight=ew, **KW) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            central, contextual, losses = engine.fit(42 + trial)
        assert losses[-1] < losses[0]
        ours = np.hstack([central, contextual])
        a_ref = auroc(reference, train_pos, test_pos, train_neg, test_neg)
        a_gpu = auroc(ours, train_pos, test_pos, train_neg, test_neg)
        print(f"{model} rw={rw} ew={ew} holdout {trial}: AUROC oracle {a_ref:.4f}  gpu {a_gpu:.4f}")
        assert a_gpu > 0.85
        assert abs(a_gpu - a_ref) <= 0.01
        deltas.append(a_gpu - a_ref)
    assert abs(np.mean(deltas)) <= 0.005  # north_star: within 0.005, two-sided


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["SkipGram", "CBOW"])
def test_loss_curve_tracks_the_oracle(model):
    """Loss-curve leg (north_star, SURVEY.md 8c leg 3): the per-epoch mean pair loss of the GPU
    run (thousands of concurrent walks) against the 8-thread Hogwild oracle on the same graph,
    kwargs and seed.  Stated tolerance: 20 % on the first two epochs (where the staleness of
    concurrent updates is largest; the number of walks in flight is capped on small graphs for
    exactly this reason, see b2e_config.max_concurrent_walks), 8 % on the third, 5 % afterwards
    (measured, CBOW: 5.6 %, 14 %, 5.0 %, 3.7 %, 3.1 %, 3.0 %; SkipGram stays below 3 % from the
    third epoch on)."""
    from embiggen_b200.engine import Engine
    src, dst, n = block_model(7)
    graph = csr_from_edges(src, dst, n)
    kw = dict(KW, epochs=6)
    oracle.set_threads(8)
    try:
        _, _, expected = oracle.fit(model, graph.indptr, graph.indices, 5, kw["embedding_size"], kw["epochs"],
                                    kw["iterations"], kw["walk_length"], kw["window_size"],
                                    kw["number_of_negative_samples"], kw["learning_rate"],
                                    kw["learning_rate_decay"], return_weight=0.5, explore_weight=2.0)
    finally:
        oracle.set_threads(1)
    with Engine(model, return_weight=0.5, explore_weight=2.0, **kw) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        _, _, got = engine.fit(5)
    print(model, "oracle", np.round(expected, 4), "gpu", np.round(got, 4))
    assert got[-1] < got[1] < got[0]
    for epoch, (a, b) in enumerate(zip(expected, got)):
        assert abs(b - a) <= (0.20 if epoch < 2 else 0.08 if epoch == 2 else 0.05) * a, (epoch, a, b)


@pytest.mark.gpu
def test_replica_averaging_keeps_the_quality():
    """The data-parallel exchange step on one GPU: two replicas (two handles) train the two
    shards of every step and are averaged at the driver's sync points, exactly like two ranks;
    the averaged embedding must predict held-out edges as well as the single-replica one
    (within 0.005 AUROC, two-sided, on one holdout; measured 0.8971 vs 0.8980)."""
    import torch
    from embiggen_b200.engine import Engine, shard_chunks
    src, dst, n = block_model(3)
    train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, 3)
    graph = csr_from_edges(train_pos[0], train_pos[1], n)
    seed, world, sync_interval = 17, 2, 2
    with Engine("SkipGram", **KW) as single:
        single.load_csr(graph.indptr, graph.indices)
        c, x, _ = single.fit(seed)
    baseline = auroc(np.hstack([c, x]), train_pos, test_pos, train_neg, test_neg)

    replicas = [Engine("SkipGram", chunk_walks=2048, **KW) for _ in range(world)]
    try:
        for engine in replicas:
            engine.load_csr(graph.indptr, graph.indices)
            engine.init_tables(seed)
        for rank, engine in enumerate(replicas):
            engine.open_exchange_local(replicas, rank)

        def average():  # the exchange kernel of every "rank", bracketed by syncs like Engine.average
            for engine in replicas:
                engine.sync()
            for engine in replicas:
                engine.exchange_average()
            for engine in replicas:
                engine.sync()

        per_epoch = replicas[0].walks_per_epoch
        lr = np.float32(KW["learning_rate"])
        for epoch in range(KW["epochs"]):
            plans = [list(shard_chunks(per_epoch, 2048, world, rank, epoch * per_epoch)) for rank in range(world)]
            for index in range(len(plans[0])):
                for rank, engine in enumerate(replicas):
                    first, count, stride = plans[rank][index]
                    engine.walk_chunk(seed, first, count, stride, index & 1)
                    engine.train_chunk(seed, index & 1, float(lr))
                if (index + 1) % sync_interval == 0:
                    average()
            average()
            lr = np.float32(lr * np.float32(KW["learning_rate_decay"]))
        a0, a1 = replicas[0].export_tables()
        b0, _ = replicas[1].export_tables()
        assert np.array_equal(a0, b0)  # replicas agree after the exchange
        assert replicas[0].tables_digest()["bits"] == replicas[1].tables_digest()["bits"]
    finally:
        for engine in replicas:
            engine.close()
    averaged = auroc(np.hstack([a0, a1]), train_pos, test_pos, train_neg, test_neg)
    print(f"AUROC single replica {baseline:.4f}  two averaged replicas {averaged:.4f}")
    assert abs(averaged - baseline) <= 0.005


@pytest.mark.gpu
def test_shared_negatives_auroc_against_the_per_pair_model():
    """The opt-in `shared_negatives` estimator (one set of negatives per centre)
    through the same held-out-edge protocol, against the default per-pair model on the same
    holdouts.  It is a different estimator, so the bound is the one stated for it, not the 0.005
    of the parity contract: every holdout above 0.85 and within 0.02 of the per-pair AUROC."""
    from embiggen_b200.engine import Engine
    deltas = []
    for trial in range(3):
        src, dst, n = block_model(trial)
        train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, trial)
        graph = csr_from_edges(train_pos[0], train_pos[1], n)
        scores = {}
        for shared in (False, True):
            with Engine("SkipGram", return_weight=1.0, explore_weight=1.0, shared_negatives=shared, **KW) as engine:
                engine.load_csr(graph.indptr, graph.indices)
                central, contextual, losses = engine.fit(42 + trial)
            assert losses[-1] < losses[0]
            scores[shared] = auroc(np.hstack([central, contextual]), train_pos, test_pos, train_neg, test_neg)
        print(f"holdout {trial}: AUROC per-pair negatives {scores[False]:.4f}  shared negatives {scores[True]:.4f}")
        assert scores[True] > 0.85
        deltas.append(scores[True] - scores[False])
    print("shared - per-pair, mean over the holdouts:", float(np.mean(deltas)))
    assert min(deltas) > -0.02
