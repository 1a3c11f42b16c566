"""Host-side mirror of the step after the embedding path: edge embeddings
and a perceptron edge scorer on node features that stay in HBM.

* :class:`EdgeTransformerB200` -- same constructor / ``fit`` / ``transform`` meaning as
  ``EdgeTransformer`` (/root/reference/embiggen/embedding_transformers/edge_transformer.py:
  337-361 the method table, :423-490 ``fit``, :511-600 ``transform``) for node ids with an
  aligned mapping; node / edge type features are not supported.
* :class:`PerceptronEdgePredictionB200` -- ``PerceptronEdgePrediction``
  (/root/reference/embiggen/edge_prediction/edge_prediction_ensmallen/perceptron.py:15-300):
  same kwargs and defaults, ``edge_embeddings`` and the topological ``edge_features`` (Degree,
  AdamicAdar, JaccardCoefficient, ResourceAllocationIndex, PreferentialAttachment; not
  Cooccurrence).
* :func:`binary_auroc` -- tie-aware AUROC of the scores (``express_measures.binary_auroc``,
  abstract_classifier_model.py:2073), on the host: scores are 4 bytes per edge, the embedding
  is what must not move.

Everything is computed by ``libb2e.so`` (``csrc/edge_pred.cu``); there is no CPU fallback.
"""
import ctypes
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import B2EPerceptronConfig, check
from .graph import as_csr

EDGE_METHODS = ["Hadamard", "Sum", "Average", "L1", "AbsoluteL1", "SquaredL2", "L2", "Concatenate",
                "Min", "Max", "L2Distance", "CosineSimilarity"]
EDGE_FEATURES = ["Degree", "AdamicAdar", "JaccardCoefficient", "ResourceAllocationIndex",
                 "PreferentialAttachment"]  # perceptron.py:38-46; "Cooccurrence" is not implemented
# perceptron.py:51-61 names some methods differently from the transformer
_PERCEPTRON_ALIASES = {"EuclideanDistance": "L2Distance", "Add": "Sum", "Sub": "L1",
                       "Maximum": "Max", "Minimum": "Min"}


def _method_ids(methods: Union[str, Sequence[str]], aliases: Optional[Dict[str, str]] = None) -> np.ndarray:
    if isinstance(methods, str):
        methods = [methods]
    ids = []
    for name in methods:
        if not isinstance(name, str):
            raise ValueError(f"The provided method name should be a string, but we got {type(name)} instead.")
        name = (aliases or {}).get(name, name)
        if name not in EDGE_METHODS:
            raise ValueError(f"The provided edge embedding method {name!r} is not in {EDGE_METHODS}.")
        ids.append(EDGE_METHODS.index(name))
    if not ids:
        raise ValueError("At least one edge embedding method is required.")
    return np.asarray(ids, dtype=np.uint32)


def _feature_ids(names: Optional[Union[str, Sequence[str]]]) -> np.ndarray:
    if names is None:
        names = []
    if isinstance(names, str):
        names = [names]
    ids = []
    for name in names:
        if name == "Cooccurrence":
            raise NotImplementedError("The Cooccurrence edge feature is not implemented by the B200 engine.")
        if name not in EDGE_FEATURES:
            raise ValueError(f"The provided edge feature {name!r} is not in {EDGE_FEATURES}.")
        ids.append(EDGE_FEATURES.index(name))
    return np.asarray(ids, dtype=np.uint32)


def edge_metrics(graph, sources, destinations, edge_features: Union[str, Sequence[str]],
                 device: int = 0) -> np.ndarray:
    """The topological edge features of an edge list, computed on the GPU from the graph's CSR
    (what ``graph.get_jaccard_coefficient_scores`` & co. return, graph_visualizer.py:2465,2677)."""
    indptr, indices, _ = as_csr(graph)
    ids = _feature_ids(edge_features)
    if len(ids) == 0:
        raise ValueError("At least one edge feature is required.")
    src, dst = _edges(sources, destinations)
    width = sum(2 if i == 0 else 1 for i in ids)
    out = np.empty((src.shape[0], width), dtype=np.float32)
    check(_lib.load().b2e_edge_metrics(device, indptr.ctypes.data, indices.ctypes.data, indptr.shape[0] - 1,
                                       indices.shape[0], src.ctypes.data, dst.ctypes.data, src.shape[0],
                                       ids.ctypes.data, len(ids), out.ctypes.data))
    return out


def _edges(sources, destinations):
    src = np.ascontiguousarray(sources, dtype=np.uint32)
    dst = np.ascontiguousarray(destinations, dtype=np.uint32)
    if src.ndim != 1 or src.shape != dst.shape:
        raise ValueError("sources and destinations must be one-dimensional and of the same length.")
    return src, dst


class DeviceFeatures:
    """Node features resident in HBM (``b2e_features``): a host matrix uploaded once, or a
    zero-copy view of the tables of a trained :class:`~embiggen_b200.engine.Engine`."""

    def __init__(self, features=None, device: int = 0, engine=None, table: int = 0):
        self._lib = _lib.load()
        self._handle = ctypes.c_void_p()
        self._keepalive = engine
        if engine is not None:
            check(self._lib.b2e_features_from_handle(engine._handle, table, ctypes.byref(self._handle)))
            self.shape = (engine.n, engine.config.embedding_size)
        else:
            features = np.ascontiguousarray(getattr(features, "values", features), dtype=np.float32)
            if features.ndim != 2 or features.size == 0:
                raise ValueError("The node features must be a non-empty matrix.")
            if not np.isfinite(features).all():
                raise ValueError("The node features contain NaN or infinite values.")
            check(self._lib.b2e_features_create(device, features.ctypes.data, features.shape[0],
                                                features.shape[1], ctypes.byref(self._handle)))
            self.shape = features.shape

    def close(self) -> None:
        if self._handle:
            self._lib.b2e_features_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_device_features(node_feature, device: int = 0):
    if isinstance(node_feature, DeviceFeatures):
        return node_feature, False
    if isinstance(node_feature, (list, tuple)):  # several embeddings: concatenated, like NodeTransformer
        node_feature = np.hstack([np.asarray(getattr(f, "values", f), dtype=np.float32) for f in node_feature])
    return DeviceFeatures(node_feature, device=device), True


class EdgeTransformerB200:
    """Counterpart of ``EdgeTransformer`` for aligned node ids."""

    methods = {name: index for index, name in enumerate(EDGE_METHODS)}

    def __init__(self, methods: Union[List[str], str] = "Hadamard", aligned_mapping: bool = True,
                 device: int = 0):
        if not aligned_mapping:
            raise NotImplementedError("Only aligned_mapping=True (node ids, not names) is supported.")
        self._method_ids = _method_ids(methods)
        self._device = device
        self._features: Optional[DeviceFeatures] = None
        self._owned = False

    def fit(self, node_feature, node_type_feature=None, edge_type_features=None) -> None:
        if node_type_feature is not None or edge_type_features:
            raise NotImplementedError("Node type and edge type features are not supported.")
        if self._features is not None and self._owned:
            self._features.close()
        self._features, self._owned = _as_device_features(node_feature, self._device)

    def embedding_size(self) -> int:
        if self._features is None:
            raise ValueError("Transformer was not fitted yet.")
        size = ctypes.c_uint32()
        check(_lib.load().b2e_edge_embedding_size(self._features.shape[1], self._method_ids.ctypes.data,
                                                  len(self._method_ids), None, 0, ctypes.byref(size)))
        return int(size.value)

    def transform(self, sources, destinations) -> np.ndarray:
        if self._features is None:
            raise ValueError("Transformer was not fitted yet.")
        src, dst = _edges(sources, destinations)
        out = np.empty((src.shape[0], self.embedding_size()), dtype=np.float32)
        check(_lib.load().b2e_edge_embedding(self._features._handle, src.ctypes.data, dst.ctypes.data,
                                             src.shape[0], self._method_ids.ctypes.data,
                                             len(self._method_ids), out.ctypes.data))
        return out


class PerceptronEdgePredictionB200:
    """Perceptron edge scorer over edge embeddings (Adam, scale-free negative sampling)."""

    def __init__(self, edge_features: Optional[Union[str, List[str]]] = "JaccardCoefficient",
                 edge_embeddings: Optional[Union[str, List[str]]] = None,
                 cooccurrence_iterations: int = 100, cooccurrence_window_size: int = 10,
                 number_of_epochs: int = 1000, number_of_edges_per_mini_batch: int = 4096,
                 learning_rate: float = 0.001, first_order_decay_factor: float = 0.9,
                 second_order_decay_factor: float = 0.999, avoid_false_negatives: bool = False,
                 use_scale_free_distribution: bool = True, random_state: int = 42, verbose: bool = True,
                 device: int = 0):
        if isinstance(edge_features, str):
            edge_features = [edge_features]
        if isinstance(edge_embeddings, str):
            edge_embeddings = [edge_embeddings]
        if not edge_features and not edge_embeddings:
            raise ValueError("At least one edge feature or edge embedding method is required.")
        self._feature_ids = _feature_ids(edge_features)
        self._method_ids = _method_ids(edge_embeddings, _PERCEPTRON_ALIASES) if edge_embeddings \
            else np.zeros(0, dtype=np.uint32)
        self._support = None
        self._model_kwargs = dict(
            edge_features=list(edge_features) if edge_features else None,
            edge_embeddings=list(edge_embeddings) if edge_embeddings else None,
            cooccurrence_iterations=cooccurrence_iterations,
            cooccurrence_window_size=cooccurrence_window_size, number_of_epochs=number_of_epochs,
            number_of_edges_per_mini_batch=number_of_edges_per_mini_batch, learning_rate=learning_rate,
            first_order_decay_factor=first_order_decay_factor,
            second_order_decay_factor=second_order_decay_factor,
            avoid_false_negatives=avoid_false_negatives,
            use_scale_free_distribution=use_scale_free_distribution)
        self._random_state = random_state
        self._verbose = verbose
        self._device = device
        self._params: Optional[np.ndarray] = None
        self._losses: List[float] = []
        _lib.load()

    # -- identity, like perceptron.py:118-131, :283-300 --
    @classmethod
    def model_name(cls) -> str:
        return "Perceptron"

    @classmethod
    def library_name(cls) -> str:
        return "B200"

    @classmethod
    def task_name(cls) -> str:
        return "Edge Prediction"

    def parameters(self) -> Dict[str, Any]:
        return dict(random_state=self._random_state, **self._model_kwargs)

    def clone(self) -> "PerceptronEdgePredictionB200":
        return PerceptronEdgePredictionB200(**self.parameters(), device=self._device)

    @classmethod
    def smoke_test_parameters(cls) -> Dict[str, Any]:
        return dict(number_of_epochs=1, edge_features="Degree")  # perceptron.py:125-131

    def get_losses(self) -> List[float]:
        return list(self._losses)

    def get_weights(self) -> np.ndarray:
        """The weights followed by the bias."""
        if self._params is None:
            raise ValueError("The model was not fitted yet.")
        return self._params.copy()

    def _config(self) -> B2EPerceptronConfig:
        k = self._model_kwargs
        config = B2EPerceptronConfig(
            struct_size=ctypes.sizeof(B2EPerceptronConfig), n_methods=len(self._method_ids),
            n_edge_features=len(self._feature_ids),
            number_of_epochs=k["number_of_epochs"],
            number_of_edges_per_mini_batch=k["number_of_edges_per_mini_batch"],
            learning_rate=k["learning_rate"], first_order_decay_factor=k["first_order_decay_factor"],
            second_order_decay_factor=k["second_order_decay_factor"],
            avoid_false_negatives=int(bool(k["avoid_false_negatives"])),
            use_scale_free_distribution=int(bool(k["use_scale_free_distribution"])))
        for i, method in enumerate(self._method_ids):
            config.methods[i] = int(method)
        for i, feature in enumerate(self._feature_ids):
            config.edge_features[i] = int(feature)
        return config

    def _ids(self):
        return (self._method_ids.ctypes.data if len(self._method_ids) else None, len(self._method_ids),
                self._feature_ids.ctypes.data if len(self._feature_ids) else None, len(self._feature_ids))

    def fit(self, graph, node_features=None) -> "PerceptronEdgePredictionB200":
        """``graph``: anything :func:`embiggen_b200.graph.as_csr` accepts (it is also the support
        of the edge features); ``node_features``: a matrix, a list of matrices (concatenated) or
        :class:`DeviceFeatures`, needed only with ``edge_embeddings``."""
        indptr, indices, _ = as_csr(graph)
        if indices.shape[0] == 0:
            raise ValueError("The provided graph does not have any edge.")
        if len(self._method_ids) and node_features is None:
            raise ValueError("The edge embeddings need node features.")
        features, owned = (None, False) if not len(self._method_ids) else \
            _as_device_features(node_features, self._device)
        try:
            config = self._config()
            size = ctypes.c_uint32()
            lib = _lib.load()
            if features is None:  # no feature matrix to carry the device: say it
                check(lib.b2e_select_device(int(self._device)))
            check(lib.b2e_edge_embedding_size(features.shape[1] if features else 0, *self._ids(),
                                              ctypes.byref(size)))
            params = np.empty(size.value + 1, dtype=np.float32)
            losses = np.zeros(max(1, config.number_of_epochs), dtype=np.float32)
            check(lib.b2e_perceptron_fit(features._handle if features else None, indptr.ctypes.data,
                                         indices.ctypes.data, indptr.shape[0] - 1, indices.shape[0],
                                         ctypes.byref(config), int(self._random_state) & 0xFFFFFFFFFFFFFFFF,
                                         params.ctypes.data, losses.ctypes.data))
        finally:
            if owned:
                features.close()
        self._params = params
        self._support = (indptr, indices)
        self._losses = [float(x) for x in losses[:config.number_of_epochs]]
        return self

    def predict_proba(self, sources, destinations, node_features=None, support=None) -> np.ndarray:
        """Scores of an edge list; ``support`` (default: the graph given to ``fit``) is the graph
        the edge features are computed on (perceptron.py:172-215)."""
        if self._params is None:
            raise ValueError("The model was not fitted yet.")
        src, dst = _edges(sources, destinations)
        indptr, indices = self._support if support is None else as_csr(support)[:2]
        if len(self._method_ids) and node_features is None:
            raise ValueError("The edge embeddings need node features.")
        features, owned = (None, False) if not len(self._method_ids) else \
            _as_device_features(node_features, self._device)
        try:
            scores = np.empty(src.shape[0], dtype=np.float32)
            if features is None:
                check(_lib.load().b2e_select_device(int(self._device)))
            check(_lib.load().b2e_perceptron_predict(
                features._handle if features else None, indptr.ctypes.data, indices.ctypes.data,
                indptr.shape[0] - 1, indices.shape[0], src.ctypes.data, dst.ctypes.data, src.shape[0],
                *self._ids(), self._params.ctypes.data, scores.ctypes.data))
        finally:
            if owned:
                features.close()
        return scores

    def predict(self, sources, destinations, node_features=None, support=None) -> np.ndarray:
        return self.predict_proba(sources, destinations, node_features, support) > 0.5


def binary_auroc(labels, scores) -> float:
    """AUROC with ties given their average rank (Mann-Whitney U over the scores)."""
    labels = np.asarray(labels).astype(bool)
    scores = np.asarray(scores, dtype=np.float64)
    if labels.shape != scores.shape:
        raise ValueError("labels and scores must have the same shape.")
    n_pos, n_neg = int(labels.sum()), int((~labels).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError("The AUROC needs at least one positive and one negative.")
    order = np.argsort(scores, kind="stable")
    sorted_scores = scores[order]
    first = np.concatenate(([True], sorted_scores[1:] != sorted_scores[:-1]))
    group = np.cumsum(first) - 1
    starts = np.flatnonzero(first)
    ends = np.concatenate((starts[1:], [len(scores)]))
    average_rank = 0.5 * (starts + ends - 1) + 1.0
    ranks = np.empty(len(scores), dtype=np.float64)
    ranks[order] = average_rank[group]
    return float((ranks[labels].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))
