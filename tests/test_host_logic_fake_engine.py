"""The embedders' host logic without a GPU: `embiggen_b200.embedders.Engine` is replaced by a
stand-in that records what it is given and fills the caller's tables with recognisable numbers,
and the bodies of the `-m gpu` embedder tests (tests/test_gpu_embedders.py, the Walklets and GloVe
classes) run through construction, `embed_graph`, graph validation, engine keyword assembly,
output buffers / files / dtypes, DataFrame wrapping and loss bookkeeping.  Nothing here computes
an embedding -- the CUDA path is tested where there is a device; this keeps the Python around it
honest on every commit.  (The stand-in is test code: the product has no CPU path.)"""
import numpy as np
import pandas as pd
import pytest

from embiggen_b200 import embedders
from embiggen_b200.embedders import (B200_EMBEDDERS, DeepWalkGloVeB200, DeepWalkSkipGramB200, Node2VecCBOWB200,
                                     Node2VecGloVeB200, Node2VecSkipGramB200, WalkletsCBOWB200,
                                     WalkletsSkipGramB200, embed_graph)
from embiggen_b200.embedding_api import EmbeddingResult, get_available_models_for_node_embedding


class StandInEngine:
    created = []

    def __init__(self, model, **kwargs):
        self.model, self.kwargs, self.calls = model.lower(), kwargs, []
        StandInEngine.created.append(self)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.calls.append("close")

    def load_csr(self, indptr, indices, weights=None):
        self.n = indptr.shape[0] - 1
        self.calls.append(("load_csr", weights is not None))

    def load_types(self, node_types, edge_types):
        self.calls.append(("load_types", node_types is not None, edge_types is not None))

    def fit(self, seed, table0=None, table1=None):
        D = self.kwargs["embedding_size"]
        for index, table in enumerate((table0, table1)):
            assert table.shape == (self.n, D) and table.dtype == np.float32
            rows = np.arange(self.n, dtype=np.float32)[:, None]
            table[:] = (index + 1) + rows * 1e-3 + (seed % 97) * 1e-5 + self.kwargs.get("walklet_scale", 0) * 0.1
        self.calls.append(("fit", seed))
        return table0, table1, [1.0 / (epoch + 1) for epoch in range(self.kwargs["epochs"])]


@pytest.fixture(autouse=True)
def stand_in(monkeypatch):
    StandInEngine.created = []
    monkeypatch.setattr(embedders, "Engine", StandInEngine)
    monkeypatch.setattr(embedders.B200Embedder, "is_available", staticmethod(lambda: True))


def test_embedding_pipeline_smoke(small_ppi, er_graph):
    frame = get_available_models_for_node_embedding()
    frame = frame[frame.library_name == "B200"]
    assert len(frame) == 4 and frame.available.all()
    for _, row in frame.iterrows():
        for graph in (small_ppi, er_graph):
            result = embed_graph(graph, row.model_name, library_name=row.library_name, smoke_test=True, verbose=False)
            assert isinstance(result, EmbeddingResult)
            tables = result.get_all_node_embedding()
            assert len(tables) == 2
            for table in tables:
                assert isinstance(table, pd.DataFrame) and table.shape == (graph.get_number_of_nodes(), 5)
                assert list(table.index) == graph.get_node_names()
            engine = StandInEngine.created[-1]
            assert engine.kwargs["walk_length"] == 4 and engine.kwargs["window_size"] == 1 and engine.kwargs["epochs"] == 1
            assert engine.model == ("cbow" if "CBOW" in row.model_name else "skipgram")
            if "DeepWalk" in row.model_name:
                assert engine.kwargs["return_weight"] == engine.kwargs["explore_weight"] == 1.0
            else:
                assert (engine.kwargs["return_weight"], engine.kwargs["explore_weight"]) == (0.25, 4.0)
            assert engine.calls[-1] == "close"


@pytest.mark.parametrize("model", B200_EMBEDDERS)
def test_fit_transform_arrays_and_losses(small_ppi, model):
    embedder = model(embedding_size=24, epochs=2, walk_length=16, iterations=2, verbose=False)
    result = embedder.fit_transform(small_ppi, return_dataframe=False)
    central, contextual = result.get_all_node_embedding()
    assert central.shape == contextual.shape == (1064, 24) and central.dtype == np.float32 and central.flags.c_contiguous
    assert central[0, 0] < contextual[0, 0]  # table 0 first, as the engine delivers them
    assert embedder.get_losses() == [1.0, 0.5] and result.embedding_method_name == model.model_name()
    engine = StandInEngine.created[-1]
    assert engine.kwargs["shared_negatives"] is False and engine.kwargs["deterministic"] is False
    assert engine.kwargs["negative_sampling_exponent"] == 0.75 and engine.kwargs["iterations"] == 2


def test_random_state_is_read_at_fit_time(small_ppi):
    embedder = Node2VecSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1, verbose=False)
    embedder.fit_transform(small_ppi, return_dataframe=False)
    embedder.set_random_state(7)  # what normalize_node_feature does between holdouts
    embedder.fit_transform(small_ppi, return_dataframe=False)
    assert [engine.calls[-2] for engine in StandInEngine.created] == [("fit", 42), ("fit", 7)]


def test_output_paths_and_dtype(small_ppi, tmp_path):
    central_path, contextual_path = str(tmp_path / "central.npy"), str(tmp_path / "contextual")
    embedder = DeepWalkSkipGramB200(embedding_size=12, epochs=1, walk_length=8, iterations=1, verbose=False,
                                    central_nodes_embedding_path=central_path,
                                    contextual_nodes_embedding_path=contextual_path)
    central, contextual = embedder.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    assert np.array_equal(np.load(central_path), central)
    assert np.array_equal(np.load(open(contextual_path, "rb")), contextual)  # the exact path, no ".npy" appended
    half = Node2VecCBOWB200(embedding_size=12, epochs=1, walk_length=8, iterations=1, verbose=False, dtype="f16",
                            central_nodes_embedding_path=str(tmp_path / "half"))
    tables = half.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    assert all(t.dtype == np.float16 for t in tables)
    assert np.load(open(str(tmp_path / "half"), "rb")).dtype == np.float16


def test_weights_types_and_isolated_nodes_reach_the_engine(small_ppi, small_ppi_weighted):
    from embiggen_b200.graph import CSRGraph
    indptr = np.concatenate([small_ppi.indptr, [small_ppi.indptr[-1]] * 3])  # three isolated nodes
    with pytest.warns(UserWarning, match="disconnected"):
        result = DeepWalkSkipGramB200(embedding_size=8, epochs=1, walk_length=8, iterations=1, verbose=False
                                      ).fit_transform(CSRGraph(indptr, small_ppi.indices, name="with_isolated"),
                                                      return_dataframe=False)
    assert result.get_all_node_embedding()[0].shape == (1067, 8)
    DeepWalkSkipGramB200(embedding_size=8, epochs=1, verbose=False).fit_transform(
        (small_ppi.indptr, small_ppi.indices), return_dataframe=False)
    assert ("load_csr", False) in StandInEngine.created[-1].calls
    Node2VecSkipGramB200(embedding_size=8, epochs=1, verbose=False).fit_transform(small_ppi_weighted, return_dataframe=False)
    assert ("load_csr", True) in StandInEngine.created[-1].calls
    typed = Node2VecSkipGramB200(embedding_size=8, epochs=1, verbose=False, change_node_type_weight=2.0)
    typed.fit_transform(small_ppi, return_dataframe=False)  # a graph without types walks untyped
    assert ("load_types", False, False) in StandInEngine.created[-1].calls
    with_types = CSRGraph(small_ppi.indptr, small_ppi.indices, node_types=np.arange(1064) % 3,
                          edge_types=np.arange(small_ppi.indices.shape[0]) % 2)
    typed.fit_transform(with_types, return_dataframe=False)
    engine = StandInEngine.created[-1]
    assert ("load_types", True, False) in engine.calls and engine.kwargs["change_node_type_weight"] == 2.0


def test_shared_negatives_reaches_the_engine(small_ppi):
    Node2VecSkipGramB200(embedding_size=8, epochs=1, verbose=False, shared_negatives=True, window_size=4
                         ).fit_transform(small_ppi, return_dataframe=False)
    assert StandInEngine.created[-1].kwargs["shared_negatives"] is True


@pytest.mark.parametrize("cls", [WalkletsSkipGramB200, WalkletsCBOWB200])
def test_walklets_scale_by_scale_path(small_ppi_weighted, cls):
    """A weighted graph takes the scale-by-scale path (one engine per scale, window 1 on the
    sub-walks of every k-th token); the one-pass path shares a resident graph and needs a device."""
    model = cls(embedding_size=24, window_size=3, epochs=2, walk_length=32, iterations=2)
    assert model.parameters()["embedding_size"] == 24 and "verbose" not in model.parameters()
    tables = model.fit_transform(small_ppi_weighted, return_dataframe=False).get_all_node_embedding()
    assert len(tables) == 6 and all(t.shape == (1064, 8) for t in tables)
    engines = StandInEngine.created[-3:]
    assert [e.kwargs["walklet_scale"] for e in engines] == [1, 2, 3]
    assert all(e.kwargs["window_size"] == 1 and e.kwargs["embedding_size"] == 8 for e in engines)
    assert not np.array_equal(tables[0], tables[2])  # scales differ
    assert model.get_losses() == [1.0, 0.5]
    frames = model.fit_transform(small_ppi_weighted).get_all_node_embedding()
    assert list(frames[0].index) == small_ppi_weighted.get_node_names()


@pytest.mark.parametrize("cls", [Node2VecGloVeB200, DeepWalkGloVeB200])
def test_glove_classes_reach_the_engine(small_ppi, cls):
    model = cls(embedding_size=16, epochs=3, walk_length=32, window_size=3, verbose=False)
    tables = model.fit_transform(small_ppi, return_dataframe=False).get_all_node_embedding()
    engine = StandInEngine.created[-1]
    assert engine.model == "glove" and engine.kwargs["glove_alpha"] == 0.75 and engine.kwargs["iterations"] == 1
    assert engine.kwargs["number_of_negative_samples"] == 0 and len(tables) == 2 and len(model.get_losses()) == 3
