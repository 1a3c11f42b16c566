"""The embedder classes end to end on the GPU, modelled on the reference's smoke tests
(/root/reference/tests/test_node_embedding_pipelines.py:17-42: every registered model through
`embed_graph(..., smoke_test=True)`; pass criterion = an EmbeddingResult without NaN / Inf)."""
import numpy as np
import pandas as pd
import pytest

from embiggen_b200.embedders import (B200_EMBEDDERS, DeepWalkSkipGramB200, Node2VecCBOWB200,
                                     Node2VecSkipGramB200, embed_graph)
from embiggen_b200.embedding_api import EmbeddingResult, get_available_models_for_node_embedding

pytestmark = pytest.mark.gpu


def test_embedding_pipeline_smoke(small_ppi, er_graph):
    """Every registered B200 model by name, smoke-test parameters, both fixtures."""
    frame = get_available_models_for_node_embedding()
    frame = frame[frame.library_name == "B200"]
    assert len(frame) == 4 and frame.available.all()
    for _, row in frame.iterrows():
        for graph in (small_ppi, er_graph):
            result = embed_graph(graph, row.model_name, library_name=row.library_name, smoke_test=True,
                                 verbose=False)
            assert isinstance(result, EmbeddingResult)
            tables = result.get_all_node_embedding()
            assert len(tables) == 2
            for table in tables:
                assert isinstance(table, pd.DataFrame)
                assert table.shape == (graph.get_number_of_nodes(), 5)
                assert list(table.index) == graph.get_node_names()
                assert np.isfinite(table.to_numpy()).all()


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_fit_transform_arrays(small_ppi, model):
    embedder = model(embedding_size=24, epochs=2, walk_length=16, iterations=2, verbose=False)
    result = embedder.fit_transform(small_ppi, return_dataframe=False)
    central, contextual = result.get_all_node_embedding()
    for table in (central, contextual):
        assert isinstance(table, np.ndarray) and table.dtype == np.float32
        assert table.shape == (1064, 24) and table.flags.c_contiguous
        assert np.isfinite(table).all() and table.std() > 0
    losses = embedder.get_losses()
    assert len(losses) == 2 and losses[1] < losses[0]
    assert result.embedding_method_name == model.model_name()


def test_random_state_is_read_at_fit_time(small_ppi):
    embedder = Node2VecSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1, verbose=False,
                                    deterministic=True)
    first = embedder.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()[0]
    again = embedder.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()[0]
    assert np.array_equal(first, again)  # deterministic engine: same seed, same tables
    embedder.set_random_state(7)         # what normalize_node_feature does between holdouts
    other = embedder.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()[0]
    assert not np.array_equal(first, other)


def test_output_paths_and_dtype(small_ppi, tmp_path):
    central_path, contextual_path = str(tmp_path / "central.npy"), str(tmp_path / "contextual.npy")
    embedder = DeepWalkSkipGramB200(embedding_size=12, epochs=1, walk_length=8, iterations=1, verbose=False,
                                    central_nodes_embedding_path=central_path,
                                    contextual_nodes_embedding_path=contextual_path)
    central, contextual = embedder.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    assert np.array_equal(np.load(central_path), central)
    assert np.array_equal(np.load(contextual_path), contextual)
    half = Node2VecCBOWB200(embedding_size=12, epochs=1, walk_length=8, iterations=1, verbose=False, dtype="f16")
    tables = half.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    assert all(t.dtype == np.float16 for t in tables)


def test_graph_without_weights_or_with_disconnected_nodes(small_ppi):
    from embiggen_b200.graph import CSRGraph
    indptr = np.concatenate([small_ppi.indptr, [small_ppi.indptr[-1]] * 3])  # three isolated nodes
    graph = CSRGraph(indptr, small_ppi.indices, name="with_isolated")
    with pytest.warns(UserWarning, match="disconnected"):
        result = DeepWalkSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1,
                                      verbose=False).fit_transform(graph, return_dataframe=False)
    assert result.get_all_node_embedding()[0].shape == (1067, 8)
    # (indptr, indices) pairs and scipy matrices are accepted too
    pair = DeepWalkSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1, verbose=False
                                ).fit_transform((small_ppi.indptr, small_ppi.indices), return_dataframe=False)
    assert pair.get_all_node_embedding()[1].shape == (1064, 8)
