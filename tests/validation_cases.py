"""`AbstractEmbeddingModel.fit_transform`'s checks of the graph
(/root/reference/embiggen/utils/abstract_models/abstract_embedding_model.py:114-198, 229-251),
probed with a configurable duck-typed graph and a configurable model, once, for BOTH base classes:
the restatement (tests/test_embedder_api.py) and the reference's own (tests/real_embiggen_probe.py).
An outcome is the exception type plus a fragment that identifies WHICH check fired, or what came
back."""
import warnings

import numpy as np


class FakeGraph:
    """Answers the accessors the base class calls; every answer is a constructor keyword."""
    DEFAULTS = dict(nodes=5, edges=True, node_types=0, edge_types=False, weights=False, negative=False,
                    disconnected=0, sorted_by_degree=False)

    def __init__(self, **answers):
        self.a = dict(self.DEFAULTS, **answers)

    def get_name(self): return "fake"
    def get_number_of_nodes(self): return self.a["nodes"]
    def has_nodes(self): return self.a["nodes"] > 0
    def has_edges(self): return self.a["edges"]
    def has_node_types(self): return self.a["node_types"] > 0
    def get_number_of_node_types(self): return self.a["node_types"]
    def has_edge_types(self): return self.a["edge_types"]
    def has_edge_weights(self): return self.a["weights"]
    def has_negative_edge_weights(self): return self.a["negative"]
    def has_disconnected_nodes(self): return self.a["disconnected"] > 0
    def get_number_of_disconnected_nodes(self): return self.a["disconnected"]
    def has_nodes_sorted_by_decreasing_outbound_node_degree(self): return self.a["sorted_by_degree"]
    def get_node_names(self): return [str(i) for i in range(self.a["nodes"])]


FRAGMENTS = ("is empty", "sorted by decreasing", "does not have node types", "node types, but the provided one",
             "does not have edge types", "does not have edge weights", "negative edge weights",
             "does not have edges", "cowardly refuse", "does not return an Embeddingresult")


def build_models(AbstractEmbeddingModel, EmbeddingResult):
    class Plain(AbstractEmbeddingModel):  # topological, optional everything, like the walk embedders
        returns = "result"

        def __init__(self):
            super().__init__(embedding_size=3, random_state=1)

        @classmethod
        def model_name(cls): return "probe"

        @classmethod
        def library_name(cls): return "probe library"

        @classmethod
        def is_stocastic(cls): return True

        @classmethod
        def is_topological(cls): return True

        @classmethod
        def requires_nodes_sorted_by_decreasing_node_degree(cls): return False

        @classmethod
        def can_use_edge_weights(cls): return True

        @classmethod
        def requires_edge_weights(cls): return False

        def is_using_edge_weights(self): return True

        @classmethod
        def requires_positive_edge_weights(cls): return True

        @classmethod
        def can_use_node_types(cls): return False

        @classmethod
        def can_use_edge_types(cls): return False

        def _fit_transform(self, graph, return_dataframe=True):
            if self.returns != "result":
                return self.returns
            return EmbeddingResult(embedding_method_name=self.model_name(),
                                   node_embeddings=np.ones((graph.get_number_of_nodes(), 3), dtype=np.float32))

    class NeedsSortedNodes(Plain):
        @classmethod
        def requires_nodes_sorted_by_decreasing_node_degree(cls): return True

    class NeedsTypesAndWeights(AbstractEmbeddingModel):
        def __init__(self):
            super().__init__(embedding_size=3, random_state=1)

        @classmethod
        def model_name(cls): return "typed probe"

        @classmethod
        def library_name(cls): return "probe library"

        @classmethod
        def is_stocastic(cls): return True

        @classmethod
        def is_topological(cls): return False

        @classmethod
        def requires_nodes_sorted_by_decreasing_node_degree(cls): return False

        @classmethod
        def requires_edge_weights(cls): return True

        @classmethod
        def requires_positive_edge_weights(cls): return False

        @classmethod
        def requires_node_types(cls): return True

        @classmethod
        def requires_edge_types(cls): return True

        def _fit_transform(self, graph, return_dataframe=True):
            return EmbeddingResult(embedding_method_name=self.model_name(),
                                   node_embeddings=np.ones((graph.get_number_of_nodes(), 3), dtype=np.float32))

    class WrongReturn(Plain):
        returns = [1, 2, 3]

    return Plain, NeedsSortedNodes, NeedsTypesAndWeights, WrongReturn


# name -> (model index into build_models, FakeGraph answers, fit_transform keywords, expected outcome)
CASES = {
    "a valid graph": (0, dict(), dict(return_dataframe=False), "EmbeddingResult"),
    "no nodes": (0, dict(nodes=0), dict(return_dataframe=False), "ValueError: is empty"),
    "no edges, topological model": (0, dict(edges=False), dict(return_dataframe=False), "ValueError: does not have edges"),
    "no edges, model that is not topological": (2, dict(edges=False, node_types=3, edge_types=True, weights=True), dict(return_dataframe=False), "EmbeddingResult"),
    "negative weights": (0, dict(weights=True, negative=True), dict(return_dataframe=False), "ValueError: negative edge weights"),
    "negative weights, model that takes them": (2, dict(node_types=2, edge_types=True, weights=True, negative=True), dict(return_dataframe=False), "EmbeddingResult"),
    "disconnected nodes warn": (0, dict(disconnected=2), dict(return_dataframe=False), "EmbeddingResult + UserWarning"),
    "nodes not sorted by degree": (1, dict(), dict(return_dataframe=False), "ValueError: sorted by decreasing"),
    "nodes sorted by degree": (1, dict(sorted_by_degree=True), dict(return_dataframe=False), "EmbeddingResult"),
    "node types missing": (2, dict(edge_types=True, weights=True), dict(return_dataframe=False), "ValueError: does not have node types"),
    "a single node type": (2, dict(node_types=1, edge_types=True, weights=True), dict(return_dataframe=False), "ValueError: node types, but the provided one"),
    "edge types missing": (2, dict(node_types=2, weights=True), dict(return_dataframe=False), "ValueError: does not have edge types"),
    "edge weights missing": (2, dict(node_types=2, edge_types=True), dict(return_dataframe=False), "ValueError: does not have edge weights"),
    "dataframe of more than 100 M nodes": (0, dict(nodes=100_000_001), dict(return_dataframe=True), "ValueError: cowardly refuse"),
    "_fit_transform returns something else": (3, dict(), dict(return_dataframe=False), "NotImplementedError: does not return an Embeddingresult"),
}


def run_cases(AbstractEmbeddingModel, EmbeddingResult):
    models = build_models(AbstractEmbeddingModel, EmbeddingResult)
    observed = {}
    for name, (index, answers, keywords, _) in CASES.items():
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            try:
                result = models[index]().fit_transform(FakeGraph(**answers), **keywords)
                outcome = type(result).__name__
            except Exception as error:
                fragment = next((f for f in FRAGMENTS if f in str(error)), str(error)[:60])
                outcome = f"{type(error).__name__}: {fragment}"
        if any("disconnected nodes" in str(w.message) for w in caught):
            outcome += " + " + caught[0].category.__name__
        observed[name] = outcome
    return observed


def expected_outcomes():
    return {name: case[3] for name, case in CASES.items()}
