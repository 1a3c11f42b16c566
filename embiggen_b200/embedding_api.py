"""The embedder-side contract of the drop-in boundary.

When ``embiggen`` itself is importable the four B200 embedders subclass its real
``AbstractEmbeddingModel`` and return its real ``EmbeddingResult`` (so they register in
``AbstractModel.MODELS_LIBRARY`` and work with ``embed_graph`` /
``edge_prediction_evaluation`` unchanged).  Where it is not (this repository's CI, the GPU
box) the compact restatements below supply the same surface, each method citing the
reference code whose behaviour it keeps:

* ``EmbeddingResult``           /root/reference/embiggen/utils/abstract_models/embedding_result.py:11-334
* ``AbstractModel`` (registry)  /root/reference/embiggen/utils/abstract_models/abstract_model.py:27-760
* ``AbstractEmbeddingModel``    /root/reference/embiggen/utils/abstract_models/abstract_embedding_model.py:12-259
* ``normalize_kwargs``          /root/reference/embiggen/utils/normalize_kwargs.py:74-135

Only what the Node2Vec / DeepWalk SkipGram / CBOW path touches is restated; the classifier
half of ``AbstractModel`` is out of scope.
"""
import gzip
import hashlib
import inspect
import json
import os
import pickle
import warnings
from typing import Any, Dict, List, Optional, Type, Union

import numpy as np
import pandas as pd

try:  # the real thing, when the reference package and its dependencies are installed
    if os.environ.get("B2E_NO_EMBIGGEN"):  # force the restatement (tests/test_reference_own_tests.py)
        raise ImportError("B2E_NO_EMBIGGEN is set")
    from embiggen.utils.abstract_models import (  # type: ignore
        AbstractEmbeddingModel, AbstractModel, EmbeddingResult, abstract_class)
    HAVE_EMBIGGEN = True
except Exception:  # ModuleNotFoundError for embiggen, ensmallen, dict_hash, ...
    HAVE_EMBIGGEN = False

Embedding = Union[np.ndarray, pd.DataFrame]

# kwarg name -> accepted type names; the subset of the reference's normalization_schemas.json
# this path uses, plus the B200-only extras (which the reference schema would reject,
# normalize_kwargs.py:127-134, hence a private schema).
KWARG_SCHEMA = {
    "embedding_size": "int", "epochs": "int", "clipping_value": "float",
    "number_of_negative_samples": "int", "walk_length": "int", "iterations": "int",
    "window_size": "int", "return_weight": "float", "explore_weight": "float",
    "change_node_type_weight": "float", "change_edge_type_weight": "float",
    "max_neighbours": ["int", "None"], "learning_rate": ["float", "str"],
    "learning_rate_decay": "float", "central_nodes_embedding_path": ["str", "None"],
    "contextual_nodes_embedding_path": ["str", "None"], "normalize_by_degree": "bool",
    "stochastic_downsample_by_degree": "bool", "normalize_learning_rate_by_degree": "bool",
    "use_scale_free_distribution": "bool", "random_state": "int", "dtype": "str",
    "verbose": "bool", "alpha": "float",
    # B200 extras
    "negative_sampling_exponent": "float", "scale_by_sqrt_dim": "bool", "deterministic": "bool",
    "shared_negatives": "bool",
    "chunk_walks": "int", "max_concurrent_walks": "int", "sync_interval": "int", "device": ["int", "None"],
}
_TYPES = {"bool": bool, "int": int, "float": float, "str": str, "None": type(None)}


def normalize_kwargs(model, kwargs: Dict[str, Any]) -> Dict[str, Any]:
    """Coerce exotic scalar types (numpy bool_, float-typed ints from pandas) and reject
    unknown names with NotImplementedError, like normalize_kwargs.py:74-135."""
    unknown = [key for key in kwargs if key not in KWARG_SCHEMA]
    if unknown:
        raise NotImplementedError(
            f"The following parameters are not supported: {unknown}. "
            f"The model is {model.model_name()} from library {model.library_name()} "
            f"for the task {model.task_name()}.")
    for key, value in list(kwargs.items()):
        names = KWARG_SCHEMA[key]
        names = [names] if isinstance(names, str) else names
        if isinstance(value, tuple(_TYPES[name] for name in names)):
            continue  # (a Python bool passes for an int here, exactly as in the reference)
        for name in names:
            if name == "None":
                continue
            try:
                kwargs[key] = _TYPES[name](value)
                break
            except (TypeError, ValueError):
                continue
        else:
            raise TypeError(
                f"The parameter {key} has the value \"{value}\" with type {type(value)} "
                f"but the expected type is {names}. The model is {model.model_name()} from "
                f"library {model.library_name()} for the task {model.task_name()}.")
    return kwargs


if not HAVE_EMBIGGEN:

    def abstract_class(klass):
        """Marker only, as in abstract_model.py:11-13."""
        return klass

    class EmbeddingResult:
        """Container of the embeddings a model produced (embedding_result.py:11-334)."""

        _KINDS = (("node_embeddings", "node embedding"), ("edge_embeddings", "edge embedding"),
                  ("node_type_embeddings", "node type embedding"),
                  ("edge_type_embeddings", "edge type embedding"))

        def __init__(self, embedding_method_name: str, node_embeddings=None, edge_embeddings=None,
                     node_type_embeddings=None, edge_type_embeddings=None):
            given = dict(node_embeddings=node_embeddings, edge_embeddings=edge_embeddings,
                         node_type_embeddings=node_type_embeddings,
                         edge_type_embeddings=edge_type_embeddings)
            self._embedding_method_name = embedding_method_name
            for attribute, label in self._KINDS:
                embeddings = given[attribute]
                if embeddings is not None and not isinstance(embeddings, list):
                    embeddings = [embeddings]  # :41-51
                for embedding in embeddings or []:
                    self._validate(embedding, label)
                setattr(self, "_" + attribute, embeddings)
            if self.is_single_embedding():  # proxy the methods of the only embedding, :114-129
                single = self.get_single_embedding()
                for name in dir(single):
                    if name.startswith("__") or hasattr(self, name):
                        continue
                    member = getattr(single, name, None)
                    if callable(member):
                        setattr(self, name, member)

        def _validate(self, embedding, label):  # :59-106
            name = self._embedding_method_name
            if not isinstance(embedding, (np.ndarray, pd.DataFrame)):
                raise ValueError(f"One of the provided {label} computed with the {name} method is "
                                 f"neither a numpy array or a pandas DataFrame, but a "
                                 f"`{type(embedding)}` object.")
            if embedding.shape[0] == 0:
                raise ValueError(f"One of the provided {label} computed with the {name} method "
                                 "is empty.")
            if embedding.shape[0] > 1_000_000:  # too large to scan, :77-79
                return
            values = embedding.to_numpy() if isinstance(embedding, pd.DataFrame) else embedding
            if np.isnan(values).any():
                raise ValueError(f"One of the provided {label} computed with the {name} method "
                                 "contains NaN values.")
            if np.isinf(values).any():
                raise ValueError(f"One of the provided {label} computed with the {name} method "
                                 f"contains {int(np.isinf(values).sum())} infinite values.")
            if np.isclose(values, 0.0).all():
                warnings.warn(f"One of the provided {label} computed with the {name} method "
                              "contains exclusively zeros.")

        def _lists(self):
            return [getattr(self, "_" + attribute) for attribute, _ in self._KINDS]

        def number_of_embeddings(self) -> int:
            return sum(len(embeddings) for embeddings in self._lists() if embeddings is not None)

        def is_single_embedding(self) -> bool:
            return self.number_of_embeddings() == 1

        def get_single_embedding(self) -> Embedding:
            assert self.is_single_embedding()
            return next(e[0] for e in self._lists() if e is not None)

        def _all(self, attribute, label) -> List[Embedding]:
            embeddings = getattr(self, attribute)
            if embeddings is None:
                raise ValueError(f"The {label} were requested but they were not computed by the "
                                 f"{self._embedding_method_name} method.")
            return embeddings

        def _at(self, attribute, label, index) -> Embedding:
            embeddings = self._all(attribute, label)
            if index >= len(embeddings):
                raise ValueError(f"The {label} computed with the {self._embedding_method_name} "
                                 f"method are {len(embeddings)}, but you requested the embedding "
                                 f"in position {index}.")
            return embeddings[index]

        def get_all_node_embedding(self):
            return self._all("_node_embeddings", "node embedding")

        def get_all_edge_embedding(self):
            return self._all("_edge_embeddings", "edge embedding")

        def get_all_node_type_embeddings(self):
            return self._all("_node_type_embeddings", "node types embedding")

        def get_all_edge_type_embeddings(self):
            return self._all("_edge_type_embeddings", "edge types embedding")

        def get_node_embedding_from_index(self, index: int):
            return self._at("_node_embeddings", "node embedding", index)

        def get_edge_embedding_from_index(self, index: int):
            return self._at("_edge_embeddings", "edge embedding", index)

        def get_node_type_embedding_from_index(self, index: int):
            return self._at("_node_type_embeddings", "node type embedding", index)

        def get_edge_type_embedding_from_index(self, index: int):
            return self._at("_edge_type_embeddings", "edge type embedding", index)

        @property
        def embedding_method_name(self) -> str:
            return self._embedding_method_name

        def dump(self) -> Dict[str, Any]:  # :323-334
            return dict(embedding_method_name=self._embedding_method_name,
                        **{attribute: getattr(self, "_" + attribute) for attribute, _ in self._KINDS})

        @staticmethod
        def load(cached: Dict[str, Any]) -> "EmbeddingResult":
            return EmbeddingResult(**cached)

    @abstract_class
    def _declines(method) -> bool:
        """abstract_model.py:16-23: a method counts as "not implemented" when its SOURCE TEXT holds
        the raise -- which is why the capability methods of subclasses never spell it out."""
        return "raise NotImplementedError" in inspect.getsource(method)

    def _capability_family(name: str):
        """`requires_<name>` / `can_use_<name>` / `is_using_<name>` with the defaults of
        abstract_model.py:156-512: each answers from its sibling when the sibling settles the
        question (cannot use => does not require; requires => can use and is using) and declines
        otherwise, so that a subclass implements exactly the ones that carry information."""
        requires, can_use, is_using = f"requires_{name}", f"can_use_{name}", f"is_using_{name}"

        def requires_default(cls) -> bool:
            try:
                if not getattr(cls, can_use)():
                    return False
            except (NotImplementedError, RecursionError):
                pass
            raise NotImplementedError(f"The `{requires}` method must be implemented in the child classes of "
                                      f"abstract model. It was not implemented in the class {cls.__name__}.")

        def can_use_default(cls) -> bool:
            try:
                if not _declines(getattr(cls, requires)) and getattr(cls, requires)():
                    return True
            except (NotImplementedError, RecursionError):
                pass
            raise NotImplementedError(f"The `{can_use}` method must be implemented in the child classes of "
                                      f"abstract model. It was not implemented in the class {cls.__name__}.")

        def is_using_default(self) -> bool:
            try:
                if getattr(self, requires)():
                    return True
            except (NotImplementedError, RecursionError):
                pass
            raise NotImplementedError(f"The `{is_using}` method must be implemented in the child classes of "
                                      f"abstract model. It was not implemented in the class "
                                      f"{self.__class__.__name__}.")

        for function, label in ((requires_default, requires), (can_use_default, can_use), (is_using_default, is_using)):
            function.__name__ = function.__qualname__ = label
        return {requires: classmethod(requires_default), can_use: classmethod(can_use_default),
                is_using: is_using_default}

    def _declined(name: str, instance_method: bool = False):
        """A method the root class only declares (abstract_model.py: `task_involves_*`,
        `is_topological`, `task_name`, `library_name`, `model_name`, `is_stocastic`, `clone`)."""
        def method(cls_or_self):
            owner = cls_or_self.__class__.__name__ if instance_method else cls_or_self.__name__
            raise NotImplementedError(f"The `{name}` method must be implemented in the child classes of "
                                      f"abstract model. It was not implemented in the class {owner}.")
        method.__name__ = method.__qualname__ = name
        return method if instance_method else classmethod(method)

    class AbstractModel:
        """abstract_model.py:27-750: registry, parameter plumbing, the capability vocabulary and the
        constructor's cross-checks of it."""

        MODELS_LIBRARY: Dict[str, Dict[str, Dict[str, Type["AbstractModel"]]]] = {}
        # the properties the constructor cross-checks (:86-92); node_type_features has the three
        # methods as well (:419-464) but is not part of that loop
        CHECKED_CAPABILITIES = ("edge_types", "node_types", "edge_weights", "edge_type_features", "edge_features")

        def __init__(self, random_state: Optional[int] = None):
            def names() -> str:  # only ever evaluated inside an error message, as in the reference
                return f"{self.model_name()} from library {self.library_name()} and task {self.task_name()}"

            if self.is_stocastic() and random_state is None:  # :41-48
                raise ValueError(
                    "The provided model is stocastic, yet no random state was provided. Please do "
                    f"provide a random state to the model {names()}.")
            if not self.is_stocastic() and random_state is not None:  # :49-56
                raise ValueError(
                    f"The provided model is not stocastic, yet a random state of `{random_state}` was "
                    f"provided. Please do not provide a random state to the model {names()}.")

            def useless(method: str, because: str) -> ValueError:
                return ValueError(
                    f"We have found an useless method in the class {self.__class__.__name__}, implementing "
                    f"method {names()}. It does not make sense to implement the `{method}` method when the "
                    f"{because}, as it is already handled in the root abstract model class.")

            if not _declines(self.can_use_edge_weights) and not self.can_use_edge_weights() and \
                    not _declines(self.requires_positive_edge_weights):  # :58-76
                raise useless("requires_positive_edge_weights", "`can_use_edge_weights` always returns False")
            for capability in self.CHECKED_CAPABILITIES:  # :78-131
                requires, can_use, is_using = (f"{prefix}_{capability}" for prefix in ("requires", "can_use", "is_using"))
                requires_method, can_use_method = getattr(self, requires), getattr(self, can_use)
                if _declines(requires_method) and _declines(can_use_method):
                    raise ValueError(
                        f"We have found a missing method implementation in the class {self.__class__.__name__}, "
                        f"implementing method {names()}. It is strictly necessary to implement either the "
                        f"`{requires}` method or the {can_use} method in order to adhere to the model interface "
                        "and facilitate the integration with the pipelines.")
                if not _declines(requires_method) and requires_method():
                    for method in (can_use, is_using):
                        if not _declines(getattr(self, method)):
                            raise useless(method, f"`{requires}` always returns True")
                if not _declines(can_use_method) and not can_use_method():
                    for method in (requires, is_using):
                        if not _declines(getattr(self, method)):
                            raise useless(method, f"`{can_use}` always returns False")
            self._random_state = random_state

        def parameters(self) -> Dict[str, Any]:  # :146-150
            return {} if self._random_state is None else dict(random_state=self._random_state)

        @classmethod
        def smoke_test_parameters(cls) -> Dict[str, Any]:
            raise NotImplementedError(f"`smoke_test_parameters` is not implemented in {cls.__name__}.")

        def into_smoke_test(self):  # :152-154
            return self.__class__(**{**self.parameters(), **self.smoke_test_parameters()})

        def set_random_state(self, random_state: int):  # :582-589
            if not self.is_stocastic():
                raise ValueError("It does not make sense to set the random state of a "
                                 "non-stocastic model.")
            self._random_state = random_state

        def consistent_hash(self) -> str:  # :555-564 (sha256 of parameters and names)
            payload = dict(**self.parameters(), model_name=self.model_name(),
                           library_name=self.library_name(), task_name=self.task_name())
            return hashlib.sha256(json.dumps(payload, sort_keys=True, default=str).encode()).hexdigest()

        @staticmethod
        def is_available() -> bool:
            return True

        @classmethod
        def requires_positive_edge_weights(cls) -> bool:  # :217-231
            try:
                if not cls.requires_edge_weights():
                    return False
            except (NotImplementedError, RecursionError):
                pass
            raise NotImplementedError(f"The `requires_positive_edge_weights` method must be implemented in the "
                                      f"child classes of abstract model. It was not implemented in the class {cls.__name__}.")

        locals().update(_capability_family("edge_weights"))
        locals().update(_capability_family("node_types"))
        locals().update(_capability_family("edge_types"))
        locals().update(_capability_family("edge_type_features"))
        locals().update(_capability_family("node_type_features"))
        locals().update(_capability_family("edge_features"))
        for _name in ("task_involves_edge_weights", "task_involves_topology", "is_topological",
                      "task_involves_node_types", "task_involves_edge_types", "task_name", "library_name",
                      "model_name", "is_stocastic"):
            locals()[_name] = _declined(_name)
        clone = _declined("clone", instance_method=True)
        del _name

        @staticmethod
        def register(model_class):  # :721-749
            by_model = AbstractModel.MODELS_LIBRARY.setdefault(model_class.task_name(), {})
            by_library = by_model.setdefault(model_class.model_name(), {})
            by_library.setdefault(model_class.library_name(), model_class)

        @staticmethod
        def get_task_data(model_name: str, task_name: str):
            if not model_name:
                raise ValueError("The provided model name is empty.")
            if not task_name:
                raise ValueError("The provided task name is empty.")
            library = AbstractModel.MODELS_LIBRARY
            if task_name not in library:
                raise ValueError(f"The provided task name {task_name!r} is not in {sorted(library)}.")
            if model_name not in library[task_name]:
                raise ValueError(f"The provided model name {model_name!r} is not in "
                                 f"{sorted(library[task_name])}.")
            return library[task_name][model_name]

        @classmethod
        def get_model_from_library(cls, model_name: str, task_name: Optional[str] = None,
                                   library_name: Optional[str] = None):  # :626-700
            if task_name is None:  # :654-669: the class's own task, else the first registered under that name
                try:
                    task_name = cls.task_name()
                except NotImplementedError as exception:
                    tasks = [task for task, models in AbstractModel.MODELS_LIBRARY.items() if model_name in models]
                    if not tasks:
                        raise ValueError(
                            f"The requested model `{model_name}` is not available. Please do provide a "
                            "valid model name to resolve this ambiguity.") from exception
                    task_name = tasks[0]
            task_data = AbstractModel.get_task_data(model_name, task_name)
            if library_name is None:
                names = list(task_data)
                if len(names) == 1:
                    library_name = names[0]
                elif "Ensmallen" in names:
                    library_name = "Ensmallen"
                else:
                    raise ValueError(
                        f"The requested model `{model_name}` is available for multiple libraries "
                        f"({names}) and no specific library was requested.")
            if library_name not in task_data:
                raise ValueError(f"The provided library name {library_name!r} is not in "
                                 f"{sorted(task_data)}.")
            model_class = task_data[library_name]
            if not model_class.is_available():
                model_class()  # surfaces its helpful error, :692-696
            return model_class

        @staticmethod
        def find_available_models(model_name: str, task_name: str):
            return [model for model in AbstractModel.get_task_data(model_name, task_name).values()
                    if model.is_available()]

    @abstract_class
    class AbstractEmbeddingModel(AbstractModel):
        """fit_transform with the reference's graph validation (abstract_embedding_model.py)."""

        def __init__(self, embedding_size: Optional[int] = None, enable_cache: bool = False,
                     ring_bell: bool = False, random_state: Optional[int] = None):
            super().__init__(random_state=random_state)
            if embedding_size is not None and not isinstance(embedding_size, int) or embedding_size == 0:
                raise ValueError("The embedding size, if provided, should be a strictly positive "
                                 f"integer but {embedding_size} was provided.")  # :37-41
            self._embedding_size = embedding_size
            self._enable_cache = enable_cache  # see _cache_path: the restatement of the @Cache at :91-95
            self._ring_bell = None

        def parameters(self) -> Dict[str, Any]:
            extra = {} if self._embedding_size is None else dict(embedding_size=self._embedding_size)
            return dict(**super().parameters(), **extra)

        @classmethod
        def requires_nodes_sorted_by_decreasing_node_degree(cls) -> bool:  # :56-62
            raise NotImplementedError("The `requires_nodes_sorted_by_decreasing_node_degree` method must be "
                                      "implemented in the child classes of abstract model.")

        @classmethod
        def get_minimum_required_number_of_node_types(cls) -> int:  # :64-67
            return 0

        def _fit_transform(self, graph, return_dataframe: bool = True):  # :69-89
            raise NotImplementedError("The `_fit_transform` method must be implemented in the child classes "
                                      "of abstract model.")

        @classmethod
        def can_use_edge_type_features(cls) -> bool:
            return False

        @classmethod
        def can_use_edge_features(cls) -> bool:
            return False

        def _validate_graph(self, graph) -> None:  # :114-180
            name = graph.get_name()
            if not graph.has_nodes():
                raise ValueError(f"The provided graph {name} is empty.")
            if self.requires_nodes_sorted_by_decreasing_node_degree() and \
                    not graph.has_nodes_sorted_by_decreasing_outbound_node_degree():  # :119-128
                raise ValueError(
                    f"The given graph {name} does not have the nodes sorted by decreasing order, therefore "
                    "the negative sampling (which follows a scale free distribution) would not approximate "
                    "well the Softmax.\nIn order to sort the given graph in such a way that the node IDs are "
                    "sorted by decreasing outbound node degrees, you can use the Graph method "
                    "`graph.sort_by_decreasing_outbound_node_degree()`.")
            if self.requires_node_types() and not graph.has_node_types():
                raise ValueError(f"The provided graph {name} does not have node types, but the "
                                 f"{self.model_name()} requires node types.")
            if self.requires_node_types() and graph.get_number_of_node_types() <= 1:  # :136-142
                raise ValueError(
                    f"The {self.model_name()} requires the graph to have at least "
                    f"{self.get_minimum_required_number_of_node_types()} node types, but the provided one "
                    f"has {graph.get_number_of_node_types()} node types.")
            if self.requires_edge_types() and not graph.has_edge_types():
                raise ValueError(f"The provided graph {name} does not have edge types, but the "
                                 f"{self.model_name()} requires edge types.")
            if self.requires_edge_weights() and not graph.has_edge_weights():
                raise ValueError(f"The provided graph {name} does not have edge weights, but the "
                                 f"{self.model_name()} requires edge weights.")
            if (self.requires_positive_edge_weights() and graph.has_edge_weights()
                    and graph.has_negative_edge_weights()):
                raise ValueError(f"The provided graph {name} has negative edge weights, but the "
                                 f"{self.model_name()} requires strictly positive edge weights.")
            if self.is_topological():
                if not graph.has_edges():
                    raise ValueError(f"The provided graph {name} does not have edges.")
                if graph.has_disconnected_nodes():
                    warnings.warn(
                        f"Please be advised that the {name} graph contains "
                        f"{graph.get_number_of_disconnected_nodes()} disconnected nodes. Node "
                        "embedding algorithms that only use topological information such as CBOW "
                        "and SkipGram are not able to provide meaningful embeddings for these nodes.")

        def _cache_path(self, graph, return_dataframe: bool) -> str:
            """`enable_cache` (abstract_embedding_model.py:91-95: cache_decorator's
            "{cache_dir}/{model_name}/{library_name}/{graph name}/{_hash}.pkl.gz", cache_dir =
            "embedding", overridable with the CACHE_DIR environment variable as there).  The hash
            covers the parameters (seed included) and the graph's shape and contents."""
            digest = hashlib.sha256()
            digest.update(repr(sorted(self.parameters().items())).encode())
            digest.update(repr((return_dataframe, graph.get_number_of_nodes())).encode())
            for getter in ("get_cumulative_node_degrees", "get_directed_destination_node_ids"):
                if hasattr(graph, getter):
                    array = np.ascontiguousarray(getattr(graph, getter)())
                    digest.update(repr(array.shape).encode())
                    digest.update(array[:: max(1, array.shape[0] // (1 << 22))].tobytes())  # <= 4 M samples
            if hasattr(graph, "has_edge_weights") and graph.has_edge_weights():
                weights = np.ascontiguousarray(graph.get_directed_edge_weights())
                digest.update(weights[:: max(1, weights.shape[0] // (1 << 22))].tobytes())
            directory = os.path.join(os.environ.get("CACHE_DIR", "embedding"), self.model_name(),
                                     self.library_name(), str(graph.get_name()).replace(os.sep, "_"))
            return os.path.join(directory, digest.hexdigest()[:32] + ".pkl.gz")

        def fit_transform(self, graph, repository: Optional[str] = None,
                          version: Optional[str] = None, return_dataframe: bool = True):  # :200-251
            if isinstance(graph, str):
                raise ValueError("Graph retrieval by name needs the `ensmallen` package; pass a "
                                 "graph object (ensmallen.Graph, CSRGraph, (indptr, indices)).")
            from .graph import as_graph
            graph = as_graph(graph)
            if return_dataframe and graph.get_number_of_nodes() > 100_000_000:
                raise ValueError(
                    "We cowardly refuse to execute this embedding with the added requirement to "
                    f"also return the dataframe version of this graph. This graph has "
                    f"{graph.get_number_of_nodes()}, and creating a Dataframe would most likely "
                    "cause an OOM on your system.")
            self._validate_graph(graph)
            cache_path = self._cache_path(graph, return_dataframe) if self._enable_cache else None
            if cache_path is not None and os.path.exists(cache_path):
                with gzip.open(cache_path, "rb") as handle:
                    return EmbeddingResult.load(pickle.load(handle))
            result = self._fit_transform(graph=graph, return_dataframe=return_dataframe)
            if cache_path is not None and isinstance(result, EmbeddingResult):
                os.makedirs(os.path.dirname(cache_path), exist_ok=True)
                temporary = f"{cache_path}.{os.getpid()}.tmp"
                with gzip.open(temporary, "wb", compresslevel=1) as handle:
                    pickle.dump(result.dump(), handle, protocol=pickle.HIGHEST_PROTOCOL)
                os.replace(temporary, cache_path)
            if not isinstance(result, EmbeddingResult):
                raise NotImplementedError(
                    f"The embedding result produced by the {self.model_name()} method from the "
                    f"library {self.library_name()} implemented in the class called "
                    f"{self.__class__.__name__} does not return an Embeddingresult but returns "
                    f"an object of type {type(result)}.")
            return result


def get_model_metadata(model_class) -> Dict[str, Any]:
    """One registry row (abstract_model.py:762-793): names, availability and the capability answers
    with `can_use_X` widened by `requires_X`."""
    try:
        row = dict(model_name=model_class.model_name(), task_name=model_class.task_name(),
                   library_name=model_class.library_name(), available=model_class.is_available())
        for capability in ("node_types", "edge_types", "edge_type_features", "edge_features", "edge_weights"):
            required = getattr(model_class, f"requires_{capability}")()
            row[f"requires_{capability}"] = required
            row[f"can_use_{capability}"] = required or getattr(model_class, f"can_use_{capability}")()
        row["requires_positive_edge_weights"] = model_class.requires_positive_edge_weights()
        return row
    except NotImplementedError as exception:
        raise NotImplementedError(
            f"Some of the mandatory static methods were not implemented in model class "
            f"{model_class.__name__}. The previous exception was: {exception}") from exception


def get_models_dataframe() -> pd.DataFrame:
    """Every registered model, available or not (abstract_model.py:796-805)."""
    return pd.DataFrame([get_model_metadata(model_class)
                         for tasks in AbstractModel.MODELS_LIBRARY.values()
                         for libraries in tasks.values()
                         for model_class in libraries.values()])


def get_available_models_for_node_embedding() -> pd.DataFrame:
    """abstract_model.py:808-811: the node-embedding models that can run HERE (for the B200 library:
    libb2e.so built and a Blackwell device visible), as the reference's tests iterate them
    (tests/test_node_embedding_pipelines.py:19)."""
    frame = get_models_dataframe()
    if frame.empty:
        return frame
    return frame[(frame.task_name == "Node Embedding") & frame.available]
