"""`embed_graph` (/root/reference/embiggen/embedders/graph_embedding_pipeline.py:10-106): what it
accepts, converts and re-raises, probed with the duck-typed graph and models of
tests/validation_cases.py, once, for BOTH implementations -- ours (embiggen_b200.embedders.embed_graph,
tests/test_embedder_api.py) and the reference's (tests/real_embiggen_probe.py)."""
import validation_cases

FRAGMENTS = ("unclear what to do", "not an embedding model", "smoke test version", "trying to compute a node embedding",
             "not in", "is not available")


def run_cases(embed_graph, AbstractEmbeddingModel, EmbeddingResult, graph_base=object):
    Plain, NeedsSortedNodes, _, WrongReturn = validation_cases.build_models(AbstractEmbeddingModel, EmbeddingResult)

    class Graph(validation_cases.FakeGraph, graph_base):
        sorted_calls = 0

        def sort_by_decreasing_outbound_node_degree(self):
            Graph.sorted_calls += 1
            return Graph(**dict(self.a, sorted_by_degree=True))

    class BadSmoke(Plain):
        @classmethod
        def smoke_test_parameters(cls):
            return dict(invalid_parameter=5)  # like the reference's own test class

    class GoodSmoke(Plain):
        converted = 0

        def __init__(self, marker=None):
            super().__init__()
            GoodSmoke.converted += marker is not None

        def parameters(self):
            return {}

        @classmethod
        def smoke_test_parameters(cls):
            return dict(marker=1)

    def outcome(call):
        try:
            return type(call()).__name__
        except Exception as error:
            fragment = next((f for f in FRAGMENTS if f in str(error)), str(error)[:50])
            return f"{type(error).__name__}: {fragment}"

    observed = {
        "a model instance": outcome(lambda: embed_graph(Graph(), Plain(), return_dataframe=False)),
        "a model instance and constructor kwargs": outcome(lambda: embed_graph(Graph(), Plain(), embedding_size=4)),
        "a class that is not a model": outcome(lambda: embed_graph(Graph(), int)),
        "something that is not a model": outcome(lambda: embed_graph(Graph(), 3.5)),
        "an unknown model name": outcome(lambda: embed_graph(Graph(), "No Such Model", library_name="B200")),
        "smoke test with invalid smoke parameters": outcome(lambda: embed_graph(Graph(), BadSmoke(), smoke_test=True)),
        "smoke test converts the model": outcome(lambda: embed_graph(Graph(), GoodSmoke(), smoke_test=True, return_dataframe=False)),
        "a failing fit is re-raised as ValueError": outcome(lambda: embed_graph(Graph(edges=False), Plain())),
        "a wrong result type is re-raised as ValueError": outcome(lambda: embed_graph(Graph(), WrongReturn(), return_dataframe=False)),
        "nodes are sorted for a model that needs it": outcome(lambda: embed_graph(Graph(), NeedsSortedNodes(), return_dataframe=False)),
    }
    observed["smoke conversions"] = GoodSmoke.converted
    observed["sort calls"] = Graph.sorted_calls
    return observed


EXPECTED = {
    "a model instance": "EmbeddingResult",
    "a model instance and constructor kwargs": "ValueError: unclear what to do",
    "a class that is not a model": "ValueError: not an embedding model",
    "something that is not a model": "ValueError: not an embedding model",
    "an unknown model name": "ValueError: is not available",
    "smoke test with invalid smoke parameters": "ValueError: smoke test version",
    "smoke test converts the model": "EmbeddingResult",
    "a failing fit is re-raised as ValueError": "ValueError: trying to compute a node embedding",
    "a wrong result type is re-raised as ValueError": "ValueError: trying to compute a node embedding",
    "nodes are sorted for a model that needs it": "EmbeddingResult",
    "smoke conversions": 1,
    "sort calls": 1,
}
