"""Host-side mirror of the reference's embedder API (CPU only; no kernel is launched).

Modelled on the reference's own tests of this surface:
/root/reference/tests/test_node_embedding_pipelines.py:83-105 (constructor round-trip),
/root/reference/tests/test_normalize_kwargs.py:10-31 (kwargs coercion),
/root/reference/tests/test_embedding_result.py:12-89 (container semantics),
/root/reference/tests/test_embed_graph_pipeline.py:54-94 (argument validation of embed_graph),
/root/reference/tests/test_abstract_model.py:123-137 (registry).
"""
import inspect

import numpy as np
import pandas as pd
import pytest

from embiggen_b200 import embedders
from embiggen_b200.embedders import (B200_EMBEDDERS, DeepWalkCBOWB200, DeepWalkSkipGramB200,
                                     Node2VecCBOWB200, Node2VecSkipGramB200, embed_graph)
from embiggen_b200.embedding_api import (AbstractEmbeddingModel, AbstractModel, EmbeddingResult,
                                         get_available_models_for_node_embedding, get_models_dataframe,
                                         normalize_kwargs)

# defaults of node2vec_skipgram.py:9-35 (Node2Vec) and deepwalk_skipgram.py:9-31 (DeepWalk)
REFERENCE_DEFAULTS = dict(
    embedding_size=100, epochs=30, clipping_value=6.0, number_of_negative_samples=10,
    walk_length=128, iterations=10, window_size=5, max_neighbours=100, learning_rate=0.01,
    learning_rate_decay=0.9, central_nodes_embedding_path=None,
    contextual_nodes_embedding_path=None, normalize_by_degree=False,
    stochastic_downsample_by_degree=False, normalize_learning_rate_by_degree=False,
    use_scale_free_distribution=True, random_state=42, dtype="f32", verbose=True)


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_signature_matches_reference_defaults(model):
    parameters = inspect.signature(model.__init__).parameters
    for name, default in REFERENCE_DEFAULTS.items():
        assert parameters[name].default == default, name
    for name in ("ring_bell", "enable_cache"):
        assert parameters[name].default is False
    if "Node2Vec" in model.model_name():
        assert parameters["return_weight"].default == 0.25
        assert parameters["explore_weight"].default == 4.0
        assert parameters["change_node_type_weight"].default == 1.0
        assert parameters["change_edge_type_weight"].default == 1.0
    else:
        assert "return_weight" not in parameters and "explore_weight" not in parameters


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_model_recreation_round_trip(model):
    """M(**M().parameters()) must construct and report the same parameters."""
    first = model()
    parameters = first.parameters()
    second = model(**parameters)
    for key, value in second.parameters().items():
        assert parameters[key] == value
    hidden = {"change_node_type_weight", "change_edge_type_weight", "alpha"}
    if "DeepWalk" in model.model_name():
        hidden |= {"return_weight", "explore_weight"}
    assert not hidden & set(parameters)
    assert parameters["embedding_size"] == 100 and parameters["random_state"] == 42


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_normalized_kwargs_construct(model):
    m = model()
    assert model(**normalize_kwargs(m, m.parameters())).parameters() == m.parameters()
    smoke = normalize_kwargs(m, dict(m.smoke_test_parameters()))
    assert model(**smoke).parameters()["walk_length"] == 4


def test_kwarg_coercion_and_rejection():
    m = Node2VecSkipGramB200(embedding_size=np.int64(16), epochs=3.0, use_scale_free_distribution=np.bool_(False),
                             learning_rate=np.float32(0.5), return_weight=2)
    p = m.parameters()
    assert p["embedding_size"] == 16 and isinstance(p["embedding_size"], int)
    assert p["epochs"] == 3 and isinstance(p["epochs"], int)
    assert p["use_scale_free_distribution"] is False
    assert isinstance(p["learning_rate"], float) and isinstance(p["return_weight"], float)
    with pytest.raises(NotImplementedError):
        Node2VecSkipGramB200(not_a_parameter=1)
    with pytest.raises(TypeError):
        Node2VecSkipGramB200(epochs="many")
    with pytest.raises(ValueError):
        Node2VecSkipGramB200(embedding_size=0)
    with pytest.raises(ValueError):
        Node2VecSkipGramB200(dtype="f8")
    typed = Node2VecSkipGramB200(change_node_type_weight=2.0, change_edge_type_weight=0.5)
    assert typed.is_using_node_types() and typed.is_using_edge_types()
    assert not Node2VecSkipGramB200().is_using_node_types() and not Node2VecSkipGramB200().is_using_edge_types()
    assert Node2VecSkipGramB200.can_use_node_types() and Node2VecSkipGramB200.can_use_edge_types()
    for invalid in (dict(change_node_type_weight=0.0), dict(change_edge_type_weight=-1.0)):
        with pytest.raises(ValueError):
            Node2VecSkipGramB200(**invalid)


def test_restated_abstract_model_cross_checks_its_capability_methods():
    """abstract_model.py:32-133 and the requires_ / can_use_ / is_using_ defaults (:156-512), probed
    with classes that break one rule each.  tests/test_real_embiggen_base.py runs the same cases
    under the reference's own class and holds it to the same table."""
    import capability_cases
    from embiggen_b200.embedding_api import AbstractModel
    assert capability_cases.run_cases(AbstractModel) == capability_cases.expected_outcomes()


def test_restated_fit_transform_checks_the_graph_like_the_reference():
    """abstract_embedding_model.py:114-198, 229-251: which check fires for which graph, in which
    order, with which exception; the same table is asserted for the reference's own class in
    tests/test_real_embiggen_base.py."""
    import validation_cases
    from embiggen_b200.embedding_api import AbstractEmbeddingModel, EmbeddingResult
    assert validation_cases.run_cases(AbstractEmbeddingModel, EmbeddingResult) == validation_cases.expected_outcomes()


def test_restated_embedding_result_behaves_like_the_reference():
    """embedding_result.py:11-334, the cases its own test file leaves out; same table asserted for
    the reference's class in tests/test_real_embiggen_base.py."""
    import embedding_result_cases
    from embiggen_b200.embedding_api import EmbeddingResult
    assert embedding_result_cases.run_cases(EmbeddingResult) == embedding_result_cases.EXPECTED


def test_embed_graph_accepts_converts_and_re_raises_like_the_reference():
    """graph_embedding_pipeline.py:10-106; the same table holds for the reference's own function
    (tests/test_real_embiggen_base.py)."""
    import embed_graph_cases
    from embiggen_b200.embedding_api import AbstractEmbeddingModel, EmbeddingResult
    assert embed_graph_cases.run_cases(embed_graph, AbstractEmbeddingModel, EmbeddingResult) == embed_graph_cases.EXPECTED


def test_shared_negatives_is_an_opt_in_skipgram_keyword():
    """B200 extra: off by default, survives the parameters() round trip and the smoke
    conversion, refused at construction where the kernel does not apply."""
    from embiggen_b200.embedders import Node2VecCBOWB200, DeepWalkSkipGramB200, Node2VecGloVeB200
    assert Node2VecSkipGramB200().parameters()["shared_negatives"] is False
    m = DeepWalkSkipGramB200(shared_negatives=True)
    assert m.parameters()["shared_negatives"] is True
    assert DeepWalkSkipGramB200(**m.parameters()).parameters() == m.parameters()
    assert m.into_smoke_test().parameters()["shared_negatives"] is True
    for model, kwargs in ((Node2VecCBOWB200, {}), (Node2VecGloVeB200, {}), (Node2VecSkipGramB200, dict(window_size=8)),
                          (Node2VecSkipGramB200, dict(number_of_negative_samples=16)),
                          (Node2VecSkipGramB200, dict(embedding_size=200))):
        with pytest.raises(ValueError, match="shared_negatives"):
            model(shared_negatives=True, **kwargs)


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_smoke_test_conversion_and_random_state(model):
    m = model(epochs=7, walk_length=64)
    smoke = m.into_smoke_test()
    assert type(smoke) is model
    p = smoke.parameters()
    assert (p["epochs"], p["embedding_size"], p["window_size"], p["walk_length"], p["max_neighbours"]) \
        == (1, 5, 1, 4, 10)
    m.set_random_state(1234)
    assert m.parameters()["random_state"] == 1234
    assert m.consistent_hash() != model(epochs=7, walk_length=64).consistent_hash()
    assert model().consistent_hash() == model().consistent_hash()


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_identity_and_capability_flags(model):
    assert model.task_name() == "Node Embedding" and model.library_name() == "B200"
    assert model.model_name() in ("Node2Vec SkipGram", "Node2Vec CBOW", "DeepWalk SkipGram", "DeepWalk CBOW")
    assert model.is_stocastic() and model.is_topological()
    assert not model.requires_nodes_sorted_by_decreasing_node_degree()
    assert not model.requires_edge_weights() and model.requires_positive_edge_weights()
    assert model.can_use_edge_weights() and model().is_using_edge_weights()
    assert not model.requires_node_types() and not model.requires_edge_types()
    assert not model.can_use_edge_type_features() and not model.can_use_edge_features()
    assert isinstance(model.is_available(), bool)
    # "implemented" is decided by inspect.getsource in the reference (abstract_model.py:16-23):
    # no capability method of ours may contain that literal
    for name in ("requires_edge_weights", "requires_positive_edge_weights", "can_use_edge_weights",
                 "is_using_edge_weights", "can_use_node_types", "can_use_edge_types", "is_stocastic"):
        assert "raise NotImplementedError" not in inspect.getsource(getattr(embedders.Node2VecB200, name))


def test_registry(monkeypatch):
    frame = get_models_dataframe()
    # the "available" view is the same frame filtered (abstract_model.py:808-811)
    monkeypatch.setattr(embedders.B200Embedder, "is_available", staticmethod(lambda: False))
    assert get_available_models_for_node_embedding().empty
    monkeypatch.setattr(embedders.B200Embedder, "is_available", staticmethod(lambda: True))
    assert sorted(get_available_models_for_node_embedding().model_name) == sorted(frame.model_name)
    ours = frame[frame.library_name == "B200"]
    assert sorted(ours.model_name) == ["DeepWalk CBOW", "DeepWalk SkipGram", "Node2Vec CBOW", "Node2Vec SkipGram"]
    assert not ours.requires_edge_weights.any()
    # Walklets are deliberately not registered (reference registry counts, test_abstract_model.py:125-126)
    assert not any("Walklets" in name for name in frame.model_name)
    for model in B200_EMBEDDERS:
        if model.is_available():
            assert AbstractEmbeddingModel.get_model_from_library(
                model.model_name(), task_name="Node Embedding", library_name="B200") is model
    with pytest.raises(ValueError):
        AbstractModel.get_task_data("No Such Model", "Node Embedding")
    with pytest.raises(ValueError):
        AbstractModel.get_task_data("", "Node Embedding")


def test_embed_graph_argument_validation(small_ppi):
    with pytest.raises(ValueError):  # kwargs together with a model instance
        embed_graph(small_ppi, Node2VecSkipGramB200(), epochs=1)
    with pytest.raises(ValueError):  # not an embedding model
        embed_graph(small_ppi, object())
    with pytest.raises(ValueError):  # unknown model name
        embed_graph(small_ppi, "HOPE")
    with pytest.raises((ValueError, TypeError)):
        embed_graph(embedding_size=5)


def test_fit_transform_validates_the_graph_before_touching_the_gpu():
    from embiggen_b200.graph import CSRGraph
    model = DeepWalkSkipGramB200()
    no_edges = CSRGraph(np.zeros(4, dtype=np.int64), np.zeros(0, dtype=np.uint32), name="no_edges")
    with pytest.raises(ValueError, match="does not have edges"):
        model.fit_transform(no_edges)
    empty = CSRGraph(np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.uint32), name="empty")
    with pytest.raises(ValueError, match="is empty"):
        model.fit_transform(empty)
    negative = CSRGraph(np.array([0, 1, 2]), np.array([1, 0]), weights=np.array([1.0, -1.0]), name="negative")
    with pytest.raises(ValueError, match="negative edge weights"):
        model.fit_transform(negative)
    with pytest.raises(ValueError):
        model.fit_transform("Cora")  # dataset retrieval needs ensmallen


# ---- EmbeddingResult (test_embedding_result.py of the reference) ----
def test_embedding_result_container():
    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    b = pd.DataFrame(a + 1, index=list("wxyz"))
    result = EmbeddingResult("Node2Vec SkipGram", node_embeddings=[a, b])
    assert result.embedding_method_name == "Node2Vec SkipGram"
    assert result.number_of_embeddings() == 2 and not result.is_single_embedding()
    assert result.get_all_node_embedding()[0] is a
    assert result.get_node_embedding_from_index(1) is b
    with pytest.raises(ValueError):
        result.get_node_embedding_from_index(2)
    with pytest.raises(ValueError):
        result.get_all_edge_embedding()
    with pytest.raises(ValueError):
        result.get_all_node_type_embeddings()
    restored = EmbeddingResult.load(result.dump())
    assert restored.get_all_node_embedding()[1].equals(b)
    single = EmbeddingResult("X", node_embeddings=a)  # wrapped into a list
    assert single.is_single_embedding() and single.get_single_embedding() is a
    assert single.mean() == a.mean()  # proxies the methods of the only embedding


def test_embedding_result_rejects_bad_embeddings():
    with pytest.raises(ValueError):
        EmbeddingResult("X", node_embeddings=[[1.0, 2.0]])
    with pytest.raises(ValueError):
        EmbeddingResult("X", node_embeddings=np.zeros((0, 3)))
    with pytest.raises(ValueError):
        EmbeddingResult("X", node_embeddings=np.array([[np.nan, 1.0]]))
    with pytest.raises(ValueError):
        EmbeddingResult("X", node_embeddings=pd.DataFrame([[np.inf, 1.0]]))
    with pytest.warns(UserWarning):
        EmbeddingResult("X", node_embeddings=np.zeros((2, 2)))


def test_csr_hand_off_from_graph_accessors(small_ppi):
    """as_csr pulls indptr / indices out of the ensmallen.Graph accessors the reference documents
    (pecanpy_embedders/node2vec.py:144-148,161), by duck typing."""
    from embiggen_b200.graph import as_csr, as_graph, validate_csr

    class GraphLike:  # only the accessors, like a real ensmallen.Graph would offer
        def get_number_of_nodes(self): return small_ppi.get_number_of_nodes()
        def get_cumulative_node_degrees(self): return small_ppi.get_cumulative_node_degrees()
        def get_directed_destination_node_ids(self): return small_ppi.indices
        def has_edge_weights(self): return False

    indptr, indices, weights = as_csr(GraphLike())
    assert np.array_equal(indptr, small_ppi.indptr) and np.array_equal(indices, small_ppi.indices)
    assert weights is None and indptr.dtype == np.int64 and indices.dtype == np.uint32
    validate_csr(indptr, indices)
    import scipy.sparse as sp
    matrix = sp.csr_matrix((np.ones(len(indices)), indices, indptr), shape=(1064, 1064))
    i2, x2, _ = as_csr(matrix)
    assert np.array_equal(i2, indptr) and np.array_equal(x2, indices)
    assert as_graph((indptr, indices)).get_number_of_nodes() == 1064
    with pytest.raises(ValueError):
        as_csr(42)
    with pytest.raises(ValueError):
        validate_csr(np.array([0, 2]), np.array([1, 0], dtype=np.uint32))  # unsorted row


def test_walklets_classes_mirror_the_reference_surface():
    """walklets.py:7-149: per-scale size = embedding_size // window_size, parameters() reports
    the public size, the classes are not in the registry (test_abstract_model.py:125-126)."""
    from embiggen_b200.embedders import WalkletsSkipGramB200, WalkletsCBOWB200
    for cls, name in ((WalkletsSkipGramB200, "Walklets SkipGram"), (WalkletsCBOWB200, "Walklets CBOW")):
        model = cls(embedding_size=100, window_size=4)
        assert model.model_name() == name and model.library_name() == "B200"
        p = model.parameters()
        assert p["embedding_size"] == 100 and p["window_size"] == 4
        assert (p["return_weight"], p["explore_weight"], p["epochs"]) == (1.0, 1.0, 30)
        again = cls(**p)
        assert again.parameters() == p
        assert model._embedding_size == 25
        smoke = model.into_smoke_test()
        assert type(smoke) is cls
    frame = get_models_dataframe()
    assert not any("Walklets" in name for name in frame.model_name)
    with pytest.raises(NotImplementedError):
        WalkletsSkipGramB200(central_nodes_embedding_path="x.npy")


def test_glove_classes_mirror_the_reference_surface():
    """node2vec_glove.py:5-140, deepwalk_glove.py: signature defaults, hidden parameters."""
    from embiggen_b200.embedders import DeepWalkGloVeB200, Node2VecGloVeB200
    m = Node2VecGloVeB200()
    p = m.parameters()
    assert (p["alpha"], p["epochs"], p["walk_length"], p["window_size"], p["learning_rate"],
            p["learning_rate_decay"], p["return_weight"], p["explore_weight"]) == \
        (0.75, 100, 512, 5, 0.05, 0.9, 0.25, 4.0)
    for hidden in ("change_node_type_weight", "change_edge_type_weight", "number_of_negative_samples",
                   "iterations"):
        assert hidden not in p
    assert Node2VecGloVeB200(**p).parameters() == p and m.model_name() == "Node2Vec GloVe"
    d = DeepWalkGloVeB200()
    q = d.parameters()
    assert q["learning_rate_decay"] == 0.99 and "return_weight" not in q
    assert DeepWalkGloVeB200(**q).parameters() == q and d.model_name() == "DeepWalk GloVe"
    assert type(d.into_smoke_test()) is DeepWalkGloVeB200
    frame = get_models_dataframe()
    assert not any("GloVe" in name for name in frame.model_name)


def test_enable_cache_stores_and_reuses_the_result(tmp_path, monkeypatch, small_ppi):
    """`enable_cache` (abstract_embedding_model.py:91-95): the second call with the same graph and
    parameters is served from "{CACHE_DIR}/{model}/{library}/{graph}/{hash}.pkl.gz"; another seed
    or another graph is another entry.  (Restatement only: with the real `embiggen` importable its
    own cache_decorator does this.)"""
    from embiggen_b200 import embedding_api
    if embedding_api.HAVE_EMBIGGEN:
        pytest.skip("the real embiggen base class brings its own cache")
    from embiggen_b200.embedders import Node2VecSkipGramB200
    monkeypatch.setenv("CACHE_DIR", str(tmp_path))
    calls = []

    def fake_fit(self, graph, return_dataframe=True):
        calls.append(self._random_state)
        n = graph.get_number_of_nodes()
        tables = [np.full((n, 4), float(self._random_state), dtype=np.float32) for _ in range(2)]
        return embedding_api.EmbeddingResult(self.model_name(), node_embeddings=tables)

    monkeypatch.setattr(Node2VecSkipGramB200, "_fit_transform", fake_fit)
    model = Node2VecSkipGramB200(embedding_size=4, enable_cache=True, random_state=5, verbose=False)
    first = model.fit_transform(small_ppi, return_dataframe=False)
    again = model.fit_transform(small_ppi, return_dataframe=False)
    assert calls == [5]
    assert np.array_equal(first.get_all_node_embedding()[1], again.get_all_node_embedding()[1])
    stored = list(tmp_path.rglob("*.pkl.gz"))
    assert len(stored) == 1 and stored[0].parts[-4:-1] == ("Node2Vec SkipGram", "B200", "small_ppi")
    model.set_random_state(6)
    model.fit_transform(small_ppi, return_dataframe=False)
    Node2VecSkipGramB200(embedding_size=4, enable_cache=False, random_state=6, verbose=False).fit_transform(
        small_ppi, return_dataframe=False)
    assert calls == [5, 6, 6] and len(list(tmp_path.rglob("*.pkl.gz"))) == 2
