"""Runs the B200 embedders under the REAL `embiggen` base classes (subprocess of
tests/test_real_embiggen_base.py; needs /root/reference, so it runs in the build container and
skips on the GPU box).

`import embiggen` pulls in a dozen third-party packages that are not installed here (ensmallen,
matplotlib, dict_hash, cache_decorator, ...).  None of them is on the path under test, so each
missing one is replaced by a permissive stub module -- discovered one ModuleNotFoundError at a
time -- except the three whose behaviour the base classes rely on: `cache_decorator.Cache` (an
identity decorator: enable_cache=False), `dict_hash` (a plain sha256) and
`userinput.utils.must_be_in_set` (case-insensitive membership, ValueError otherwise).  What then executes is
the reference's own `AbstractModel.__init__` (abstract_model.py:27-131, the inspect.getsource
"no useless method" cross-checks), `AbstractEmbeddingModel.fit_transform` /
`_cached_fit_transform` (abstract_embedding_model.py:91-251), `EmbeddingResult`
(embedding_result.py) and the registry (abstract_model.py:640-749)."""
import hashlib
import importlib.abc
import importlib.machinery
import json
import os
import sys
import types
from unittest import mock

REFERENCE = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Permissive(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        value = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, value)
        return value


class StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = {"ensmallen"}

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        module = Permissive(spec.name)
        module.__path__ = []
        return module

    def exec_module(self, module):
        pass


def install_real_base_classes():
    sys.meta_path.append(StubFinder())
    cache_decorator = Permissive("cache_decorator")
    cache_decorator.__path__ = []
    cache_decorator.Cache = lambda *args, **kwargs: (lambda function: function)
    sys.modules["cache_decorator"] = cache_decorator
    dict_hash = Permissive("dict_hash")
    dict_hash.__path__ = []
    dict_hash.Hashable = type("Hashable", (), {})
    dict_hash.sha256 = lambda obj, **_: hashlib.sha256(json.dumps(obj, sort_keys=True, default=str).encode()).hexdigest()
    sys.modules["dict_hash"] = dict_hash

    def must_be_in_set(value, candidates, label="value"):
        for candidate in candidates:
            if str(candidate).lower() == str(value).lower():
                return candidate
        raise ValueError(f"The provided {label} {value!r} is not in {sorted(map(str, candidates))}.")

    userinput = Permissive("userinput")
    userinput.__path__ = []
    userinput_utils = Permissive("userinput.utils")
    userinput_utils.__path__ = []
    userinput_utils.must_be_in_set = must_be_in_set
    userinput.utils = userinput_utils
    sys.modules["userinput"] = userinput
    sys.modules["userinput.utils"] = userinput_utils
    ensmallen = Permissive("ensmallen")  # `isinstance(graph, Graph)` needs a real class (utils/pipeline.py:44)
    ensmallen.__path__ = []
    ensmallen.Graph = type("Graph", (), {})
    sys.modules["ensmallen"] = ensmallen
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, ROOT)
    import_discovering_stubs("embiggen.utils.abstract_models")
    return sorted(StubFinder.roots)


def import_discovering_stubs(module_name):
    """Import a module of the reference, stubbing one missing third-party root per attempt."""
    import importlib
    for _ in range(80):
        try:
            return importlib.import_module(module_name)
        except ModuleNotFoundError as error:
            root = error.name.split(".")[0]
            if root == "embiggen" or root in StubFinder.roots:
                raise
            StubFinder.roots.add(root)
            # half-imported reference modules are dropped; the ones already bound by callers stay valid
            for name in [m for m in sys.modules if (m == "embiggen" or m.startswith("embiggen."))
                         and getattr(sys.modules[m], "__spec__", None) is not None
                         and getattr(sys.modules[m].__spec__, "_initializing", False)]:
                del sys.modules[name]
    raise RuntimeError("too many missing modules")


def main():
    import numpy as np
    stubs = install_real_base_classes()
    from embiggen.utils.abstract_models import AbstractEmbeddingModel, AbstractModel, EmbeddingResult
    from embiggen_b200 import embedding_api, embedders
    from embiggen_b200.graph import CSRGraph, erdos_renyi
    report = {"stubs": stubs, "have_embiggen": embedding_api.HAVE_EMBIGGEN}
    assert embedding_api.HAVE_EMBIGGEN and embedding_api.AbstractEmbeddingModel is AbstractEmbeddingModel
    assert embedding_api.EmbeddingResult is EmbeddingResult

    names = []
    for cls in embedders.B200_EMBEDDERS + (embedders.WalkletsSkipGramB200, embedders.Node2VecGloVeB200):
        model = cls()  # the reference's ctor-time checks run here (abstract_model.py:41-131)
        assert isinstance(model, AbstractEmbeddingModel) and isinstance(model, AbstractModel)
        assert model.library_name() == "B200" and model.task_name() == "Node Embedding"
        clone = cls(**model.parameters())  # tests/test_node_embedding_pipelines.py:83-105
        assert clone.parameters() == model.parameters()
        smoke = model.into_smoke_test()  # abstract_model.py:152-154
        assert smoke.parameters()["embedding_size"] == 5 or "Walklets" in model.model_name()
        model.set_random_state(7)
        assert model.parameters()["random_state"] == 7
        assert model.is_stocastic() and model.is_topological() and model.can_use_edge_weights()
        assert not model.requires_edge_weights() and model.requires_positive_edge_weights()
        assert isinstance(model.consistent_hash(), str)
        names.append(model.model_name())
    report["models"] = names

    # the capability cross-checks, probed with classes built to violate them one at a time
    # (tests/capability_cases.py): what the REFERENCE's AbstractModel does with each
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import capability_cases
    report["capability_cases"] = capability_cases.run_cases(AbstractModel)
    # fit_transform's checks of the graph (abstract_embedding_model.py:114-198, 229-251), same idea
    import validation_cases
    report["validation_cases"] = validation_cases.run_cases(AbstractEmbeddingModel, EmbeddingResult)
    import embedding_result_cases
    report["embedding_result_cases"] = embedding_result_cases.run_cases(EmbeddingResult)
    # the reference's adapter classes themselves, described as data (`ensmallen.models` is a stub:
    # their constructors run, their engines do not exist)
    import adapter_cases
    import compress_json  # a stub: hand normalize_kwargs its real schema file (normalize_kwargs.py:90)
    schema = json.load(open(os.path.join(REFERENCE, "embiggen", "utils", "normalization_schemas.json")))
    compress_json.local_load = lambda name, use_cache=True: schema
    from embiggen.embedders import ensmallen_embedders as reference_adapters
    report["adapter_description"] = adapter_cases.describe({
        "Node2Vec SkipGram": reference_adapters.Node2VecSkipGramEnsmallen,
        "Node2Vec CBOW": reference_adapters.Node2VecCBOWEnsmallen,
        "DeepWalk SkipGram": reference_adapters.DeepWalkSkipGramEnsmallen,
        "DeepWalk CBOW": reference_adapters.DeepWalkCBOWEnsmallen,
        "Walklets SkipGram": reference_adapters.WalkletsSkipGramEnsmallen,
        "Walklets CBOW": reference_adapters.WalkletsCBOWEnsmallen,
        "Node2Vec GloVe": reference_adapters.Node2VecGloVeEnsmallen,
        "DeepWalk GloVe": reference_adapters.DeepWalkGloVeEnsmallen,
    })
    PerceptronEdgePrediction = import_discovering_stubs(
        "embiggen.edge_prediction.edge_prediction_ensmallen.perceptron").PerceptronEdgePrediction
    report["perceptron_description"] = adapter_cases.describe_perceptron(PerceptronEdgePrediction)
    # the reference's embed_graph itself (graph_embedding_pipeline.py:10-106); its iterate_graphs
    # wants instances of ensmallen.Graph, so the fake graph inherits from the stub class
    import embed_graph_cases
    import ensmallen as stub_ensmallen
    from embiggen.embedders.graph_embedding_pipeline import embed_graph as real_embed_graph
    report["embed_graph_cases"] = embed_graph_cases.run_cases(real_embed_graph, AbstractEmbeddingModel, EmbeddingResult,
                                                              graph_base=stub_ensmallen.Graph)

    # the registry (abstract_model.py:640-749): the four models resolve under library "B200";
    # Walklets / GloVe were deliberately not registered
    for name in ("Node2Vec SkipGram", "Node2Vec CBOW", "DeepWalk SkipGram", "DeepWalk CBOW"):
        found = AbstractEmbeddingModel.get_model_from_library(model_name=name, task_name="Node Embedding",
                                                              library_name="B200")
        assert found.__module__ == "embiggen_b200.embedders" and found.model_name() == name
    registered = AbstractModel.MODELS_LIBRARY["Node Embedding"]
    assert "B200" not in registered.get("Walklets SkipGram", {}) and "B200" not in registered.get("Node2Vec GloVe", {})
    report["registered"] = sorted(name for name, libraries in registered.items() if "B200" in libraries)
    # the reference's own registry table (abstract_model.py:762-805) for our four classes
    from embiggen.utils.abstract_models.abstract_model import get_models_dataframe
    frame = get_models_dataframe()
    frame = frame[frame.library_name == "B200"].sort_values("model_name")
    report["registry_rows"] = json.loads(frame.to_json(orient="records"))
    try:
        AbstractEmbeddingModel.get_model_from_library(model_name="Node2Vec SkipGram", task_name="Node Embedding",
                                                      library_name="no such library")
        raise AssertionError("an unknown library must be refused")
    except ValueError:
        pass

    # the reference's own fit_transform (validation, result type check) around our _fit_transform:
    # the duck-typed graph offers every accessor the real base class calls
    graph = erdos_renyi(300, 1500, seed=3)
    calls = []

    def fake_fit(self, graph, return_dataframe=True):
        calls.append(return_dataframe)
        n = graph.get_number_of_nodes()
        tables = [np.full((n, 4), 0.5, dtype=np.float32), np.full((n, 4), 0.25, dtype=np.float32)]
        if return_dataframe:
            import pandas as pd
            tables = [pd.DataFrame(t, index=graph.get_node_names()) for t in tables]
        return EmbeddingResult(embedding_method_name=self.model_name(), node_embeddings=tables)

    original = embedders.Node2VecB200._fit_transform
    embedders.Node2VecB200._fit_transform = fake_fit
    try:
        model = embedders.Node2VecSkipGramB200(embedding_size=4, verbose=False)
        result = model.fit_transform(graph, return_dataframe=False)
        assert isinstance(result, EmbeddingResult) and len(result.get_all_node_embedding()) == 2
        framed = model.fit_transform(graph)
        assert list(framed.get_node_embedding_from_index(0).index) == graph.get_node_names()
        assert calls == [False, True]
        empty = CSRGraph(np.zeros(4, dtype=np.int64), np.zeros(0, dtype=np.uint32), name="no_edges")
        try:
            model.fit_transform(empty)
            raise AssertionError("a graph without edges must be refused")
        except ValueError as error:
            assert "does not have edges" in str(error)
    finally:
        embedders.Node2VecB200._fit_transform = original

    # without a GPU the real path fails loudly inside the engine, never on the CPU ...
    import torch
    if not torch.cuda.is_available():
        try:
            embedders.Node2VecSkipGramB200(embedding_size=4, verbose=False).fit_transform(graph)
            raise AssertionError("no CUDA device: the product path must fail")
        except RuntimeError as error:
            assert "CUDA" in str(error) or "no CPU fallback" in str(error)
        # ... and the reference's embed_graph re-wraps it as ValueError (graph_embedding_pipeline.py:99-106)
        from embiggen.embedders.graph_embedding_pipeline import embed_graph
        import ensmallen
        as_ensmallen = type("StubGraph", (CSRGraph, ensmallen.Graph), {})(graph.indptr, graph.indices, name="stub")
        try:
            embed_graph(as_ensmallen, "Node2Vec SkipGram", library_name="B200", embedding_size=4, verbose=False)
            raise AssertionError("embed_graph must re-raise")
        except ValueError:
            report["embed_graph"] = "real embed_graph resolved library B200 and re-raised as ValueError"
    print("REAL_EMBIGGEN_OK " + json.dumps(report))


if __name__ == "__main__":
    main()
