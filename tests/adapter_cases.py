"""The adapter classes of the path -- Node2Vec / DeepWalk SkipGram and CBOW, the Walklets and GloVe
siblings -- described as data: constructor signature and defaults, `parameters()` of a default and
of a customised instance, smoke parameters and smoke conversion, names and capability answers.
tests/real_embiggen_probe.py describes the REFERENCE's classes
(/root/reference/embiggen/embedders/ensmallen_embedders/{node2vec,deepwalk}_{skipgram,cbow,glove}.py,
walklets_{skipgram,cbow}.py; `ensmallen.models` stubbed, so their constructors run without the
wheel), tests/test_embedder_api.py describes ours, and the two descriptions must be equal once the
documented differences are taken out: the library name and the keyword-only B200 extras."""
import inspect

CUSTOM = dict(embedding_size=24, epochs=3, walk_length=16, window_size=2, learning_rate=0.05,
              normalize_by_degree=True, random_state=7, verbose=False)
TYPED = dict(change_node_type_weight=2.0, change_edge_type_weight=0.5)  # Node2Vec variants only


def _jsonable(value):
    if isinstance(value, dict):
        return {str(k): _jsonable(v) for k, v in sorted(value.items())}
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    if value is inspect.Parameter.empty:
        return "<required>"
    if isinstance(value, (bool, int, float, str)) or value is None:
        return value
    return repr(value)


def describe(classes, drop=()):
    """{logical name: description}; `drop` are keyword names to leave out of signatures and
    parameter dictionaries (our extras)."""
    def clean(mapping):
        return _jsonable({k: v for k, v in mapping.items() if k not in drop})

    out = {}
    for name, cls in classes.items():
        signature = inspect.signature(cls.__init__).parameters
        positional = [p for p in signature.values() if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)
                      and p.name != "self" and p.name not in drop]
        model = cls()
        typed_kwargs = {k: v for k, v in TYPED.items() if k in signature}
        custom = cls(**{k: v for k, v in CUSTOM.items() if k in signature}, **typed_kwargs)
        answers = {}
        for method in ("model_name", "task_name", "is_stocastic", "is_topological",
                       "requires_nodes_sorted_by_decreasing_node_degree", "requires_edge_weights",
                       "requires_positive_edge_weights", "can_use_edge_weights", "requires_node_types",
                       "can_use_node_types", "requires_edge_types", "can_use_edge_types",
                       "can_use_edge_type_features", "can_use_edge_features"):
            answers[method] = getattr(cls, method)()
        for method in ("is_using_edge_weights", "is_using_node_types", "is_using_edge_types"):
            answers[method] = [getattr(model, method)(), getattr(custom, method)()]
        out[name] = dict(
            signature=[[p.name, _jsonable(p.default)] for p in positional],
            parameters=clean(model.parameters()),
            custom_parameters=clean(custom.parameters()),
            smoke_test_parameters=_jsonable(cls.smoke_test_parameters()),
            smoke_converted=clean(custom.into_smoke_test().parameters()),
            recreated=clean(cls(**custom.parameters()).parameters()),
            answers=_jsonable(answers),
        )
    return out


def describe_perceptron(cls, drop=()):
    """The edge-prediction perceptron (perceptron.py:15-300) the same way."""
    signature = inspect.signature(cls.__init__).parameters
    model = cls()
    custom = cls(edge_features=["Degree", "AdamicAdar"], edge_embeddings="Hadamard", number_of_epochs=7,
                 learning_rate=0.01, random_state=3)
    return dict(
        signature=[[p.name, _jsonable(p.default)] for p in signature.values()
                   if p.name != "self" and p.name not in drop],
        parameters=_jsonable({k: v for k, v in model.parameters().items() if k not in drop}),
        custom_parameters=_jsonable({k: v for k, v in custom.parameters().items() if k not in drop}),
        smoke_test_parameters=_jsonable(cls.smoke_test_parameters()),
        names=[cls.model_name(), cls.task_name()],
    )
