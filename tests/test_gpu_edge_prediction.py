"""SURVEY.md 8(f) row 4 on the GPU, through the C ABI: edge embeddings against the reference's
own outputs (elementwise methods bit-exact, the two scalar methods to 2e-6), the perceptron
against the numpy oracle (stated tolerance: parameters 2e-3 absolute after 40 Adam steps, losses
1 %), predictions against the oracle, and the scorer on a resident embedding."""
import os

import numpy as np
import pytest

from oracle import edge_prediction as ep
from embiggen_b200.edge_prediction import (DeviceFeatures, EdgeTransformerB200,
                                           PerceptronEdgePredictionB200, binary_auroc)

pytestmark = pytest.mark.gpu
GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "edge_embedding_golden.npz"))
SCALAR = ("L2Distance", "CosineSimilarity")


@pytest.mark.parametrize("case", range(4))
def test_edge_embeddings_against_the_reference_outputs(case):
    features, src, dst = (GOLDEN[f"case{case}_{k}"] for k in ("features", "src", "dst"))
    for method in ep.METHODS:
        transformer = EdgeTransformerB200(method)
        transformer.fit(features)
        got = transformer.transform(src, dst)
        expected = GOLDEN[f"case{case}_{method}"]
        assert got.shape == expected.shape
        if method in SCALAR:
            assert np.allclose(got, expected, rtol=2e-6, atol=2e-6), method
        else:
            assert np.array_equal(got, expected), method
    # all methods at once, concatenated in the order given
    transformer = EdgeTransformerB200(list(reversed(ep.METHODS)))
    transformer.fit(features)
    got = transformer.transform(src, dst)
    expected = np.hstack([GOLDEN[f"case{case}_{m}"] for m in reversed(ep.METHODS)])
    assert got.shape == expected.shape and np.allclose(got, expected, rtol=2e-6, atol=2e-6)
    assert transformer.transform([], []).shape == (0, expected.shape[1])
    with pytest.raises(ValueError):
        transformer.transform([features.shape[0]], [0])  # node id out of range


@pytest.mark.parametrize("methods,scale_free,avoid", [
    (["Hadamard"], True, False), (["Concatenate", "CosineSimilarity"], False, True),
    (["L1", "L2Distance", "Max", "Average"], True, True)])
def test_perceptron_fit_tracks_the_oracle(small_ppi, methods, scale_free, avoid):
    rng = np.random.default_rng(3)
    features = rng.normal(size=(small_ppi.get_number_of_nodes(), 24)).astype(np.float32)
    kw = dict(number_of_epochs=2, number_of_edges_per_mini_batch=300, learning_rate=0.01)
    expected, expected_loss = ep.perceptron_fit(
        features, small_ppi.indptr, small_ppi.indices, methods, 42, 2, 300, learning_rate=0.01,
        avoid_false_negatives=avoid, scale_free=scale_free)           # 2 x 20 steps
    model = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings=methods, avoid_false_negatives=avoid,
                                         use_scale_free_distribution=scale_free, random_state=42, **kw)
    model.fit(small_ppi, features)
    got = model.get_weights()
    assert got.shape == expected.shape
    assert np.abs(got - expected).max() <= 2e-3, np.abs(got - expected).max()
    assert np.allclose(model.get_losses(), expected_loss, rtol=1e-2)
    # zero steps: the initialisation alone is bit-exact
    init = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings=methods, number_of_epochs=0,
                                        random_state=42)
    init.fit(small_ppi, features)
    assert np.array_equal(init.get_weights(), ep.perceptron_init(42, len(expected) - 1))
    src = rng.integers(0, features.shape[0], 700)
    dst = rng.integers(0, features.shape[0], 700)
    scores = model.predict_proba(src, dst, features)
    assert np.allclose(scores, ep.perceptron_predict(features, src, dst, methods, got), atol=1e-5)


def test_scorer_on_a_resident_embedding_predicts_held_out_edges():
    """Embedding trained and scored without leaving HBM: the engine's tables are viewed in place
    (b2e_features_from_handle), the perceptron is fitted on the training graph and its AUROC on
    held-out edges is compared with the same scorer fitted by the oracle on the exported table."""
    from embiggen_b200.engine import Engine
    from embiggen_b200.graph import csr_from_edges
    from test_quality import block_model, holdout
    src, dst, n = block_model(3)
    train_pos, test_pos, _, test_neg = holdout(src, dst, n, 1)
    graph = csr_from_edges(train_pos[0], train_pos[1], n)
    with Engine("SkipGram", embedding_size=32, walk_length=32, window_size=4, iterations=3, epochs=4,
                number_of_negative_samples=5, learning_rate=0.05) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        central, _, _ = engine.fit(42)
        resident = DeviceFeatures(engine=engine, table=0)
        model = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=20,
                                             number_of_edges_per_mini_batch=1024, learning_rate=0.02)
        model.fit(graph, resident)
        edges_src = np.concatenate([test_pos[0], test_neg[:, 0]])
        edges_dst = np.concatenate([test_pos[1], test_neg[:, 1]])
        scores = model.predict_proba(edges_src, edges_dst, resident)
        assert np.allclose(scores, model.predict_proba(edges_src, edges_dst, central), atol=1e-6)
        resident.close()
    labels = np.concatenate([np.ones(len(test_pos[0])), np.zeros(len(test_neg))])
    gpu_auroc = binary_auroc(labels, scores)
    params, _ = ep.perceptron_fit(central, graph.indptr, graph.indices, ["Hadamard"], 42, 20, 1024,
                                  learning_rate=0.02)
    oracle_auroc = ep.binary_auroc(labels, ep.perceptron_predict(central, edges_src, edges_dst, ["Hadamard"], params))
    print(f"perceptron AUROC on held-out edges: gpu {gpu_auroc:.4f} oracle {oracle_auroc:.4f}")
    assert gpu_auroc > 0.8 and abs(gpu_auroc - oracle_auroc) <= 0.005


def test_edge_metrics_against_the_oracle(small_ppi, rmat_graph):
    from conftest import tiny_graphs
    from embiggen_b200.edge_prediction import edge_metrics
    rng = np.random.default_rng(5)
    for graph in (small_ppi, rmat_graph, tiny_graphs()["two_components_isolated"], tiny_graphs()["star"]):
        n = graph.get_number_of_nodes()
        src, dst = rng.integers(0, n, 400), rng.integers(0, n, 400)
        rows = np.repeat(np.arange(n), np.diff(graph.indptr))
        k = min(100, graph.indices.shape[0])
        src[:k], dst[:k] = rows[:k], graph.indices[:k]            # real edges too
        for names in (ep.EDGE_FEATURES, ["JaccardCoefficient"], ["PreferentialAttachment", "Degree", "AdamicAdar"]):
            expected = ep.edge_metrics(names, graph.indptr, graph.indices, src, dst)
            got = edge_metrics(graph, src, dst, names)
            assert got.shape == expected.shape
            assert np.allclose(got, expected, rtol=1e-5, atol=1e-7), names


@pytest.mark.parametrize("edge_features,edge_embeddings", [
    ("JaccardCoefficient", None), (["Degree", "AdamicAdar", "ResourceAllocationIndex"], None),
    (["PreferentialAttachment", "JaccardCoefficient"], ["Hadamard", "L2Distance"])])
def test_perceptron_with_edge_features_tracks_the_oracle(small_ppi, edge_features, edge_embeddings):
    rng = np.random.default_rng(4)
    features = rng.normal(size=(small_ppi.get_number_of_nodes(), 16)).astype(np.float32) if edge_embeddings else None
    names = [edge_features] if isinstance(edge_features, str) else edge_features
    expected, expected_loss = ep.perceptron_fit(
        features, small_ppi.indptr, small_ppi.indices, edge_embeddings or [], 42, 2, 300, learning_rate=0.01,
        edge_features=names)
    model = PerceptronEdgePredictionB200(edge_features=edge_features, edge_embeddings=edge_embeddings,
                                         number_of_epochs=2, number_of_edges_per_mini_batch=300,
                                         learning_rate=0.01, random_state=42)
    model.fit(small_ppi, features)
    got = model.get_weights()
    assert got.shape == expected.shape and np.abs(got - expected).max() <= 2e-3
    assert np.allclose(model.get_losses(), expected_loss, rtol=1e-2)
    src = rng.integers(0, 1064, 500)
    dst = rng.integers(0, 1064, 500)
    scores = model.predict_proba(src, dst, features)
    reference = ep.perceptron_predict(features, src, dst, edge_embeddings or [], got, names,
                                      small_ppi.indptr, small_ppi.indices)
    assert np.allclose(scores, reference, atol=1e-5)


def test_default_perceptron_configuration(small_ppi):
    """The reference's default configuration (Jaccard only, no node features at all): the scores
    of edges and random pairs rank like the oracle's (AUROC within 0.005)."""
    kw = dict(number_of_epochs=10, number_of_edges_per_mini_batch=512, learning_rate=0.05)
    model = PerceptronEdgePredictionB200(**kw)
    model.fit(small_ppi)
    n = small_ppi.get_number_of_nodes()
    rows = np.repeat(np.arange(n), np.diff(small_ppi.indptr))
    rng = np.random.default_rng(0)
    src = np.concatenate([rows[::4], rng.integers(0, n, 1500)])
    dst = np.concatenate([small_ppi.indices[::4], rng.integers(0, n, 1500)])
    labels = np.concatenate([np.ones(len(rows[::4])), np.zeros(1500)])
    scores = model.predict_proba(src, dst)
    params, losses = ep.perceptron_fit(None, small_ppi.indptr, small_ppi.indices, [], 42, 10, 512,
                                       learning_rate=0.05, edge_features=["JaccardCoefficient"])
    reference = ep.perceptron_predict(None, src, dst, [], params, ["JaccardCoefficient"], small_ppi.indptr,
                                      small_ppi.indices)
    gpu_auroc, oracle_auroc = binary_auroc(labels, scores), ep.binary_auroc(labels, reference)
    print(f"default perceptron (Jaccard): AUROC gpu {gpu_auroc:.4f} oracle {oracle_auroc:.4f}")
    assert np.isfinite(scores).all() and abs(gpu_auroc - oracle_auroc) <= 0.005
    assert np.allclose(model.get_losses(), losses, rtol=1e-2)


def test_error_paths(small_ppi):
    features = np.ones((small_ppi.get_number_of_nodes() - 1, 4), dtype=np.float32)
    with pytest.raises(ValueError):
        PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=1).fit(
            small_ppi, features)  # node count mismatch
    with pytest.raises(ValueError):
        PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard").fit(small_ppi)  # no features
    with pytest.raises(ValueError):
        DeviceFeatures(np.full((3, 2), np.nan, dtype=np.float32))
    with pytest.raises(ValueError):
        PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=1,
                                     first_order_decay_factor=1.0).fit(
            small_ppi, np.ones((small_ppi.get_number_of_nodes(), 4), dtype=np.float32))
