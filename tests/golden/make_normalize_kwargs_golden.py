"""Regenerates tests/golden/normalize_kwargs_golden.json by EXECUTING the reference's own
`normalize_kwargs` (/root/reference/embiggen/utils/normalize_kwargs.py:74-135, schema
normalization_schemas.json) on exotic-typed values of every kwarg the Node2Vec / DeepWalk
SkipGram / CBOW constructors take (what pandas / JSON round trips produce: numpy scalars, floats
holding integers, ints for floats, 0/1 for bools).

    python tests/golden/make_normalize_kwargs_golden.py        # needs /root/reference

`embiggen` is imported with its uninstalled third-party dependencies stubbed
(tests/real_embiggen_probe.py); `compress_json.local_load` is pointed at the schema file."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import real_embiggen_probe as probe  # noqa: E402

CASES = {
    "int": [np.float64(5.0), np.int32(7), 3.0, np.int64(12), True],
    "float": [np.float32(0.5), 1, np.int64(2), np.float64(0.25)],
    "bool": [np.bool_(True), np.bool_(False), 0, 1],
    "str": [np.str_("f32")],
}
KEYS = ["embedding_size", "epochs", "clipping_value", "number_of_negative_samples", "walk_length", "iterations",
        "window_size", "return_weight", "explore_weight", "change_node_type_weight", "change_edge_type_weight",
        "max_neighbours", "learning_rate", "learning_rate_decay", "central_nodes_embedding_path",
        "contextual_nodes_embedding_path", "normalize_by_degree", "stochastic_downsample_by_degree",
        "normalize_learning_rate_by_degree", "use_scale_free_distribution", "random_state", "dtype", "verbose", "alpha"]


def encode(value):
    return {"type": type(value).__name__, "repr": repr(value)}


def main():
    probe.install_real_base_classes()
    import compress_json
    schema_path = os.path.join(probe.REFERENCE, "embiggen", "utils", "normalization_schemas.json")
    schema = json.load(open(schema_path))
    compress_json.local_load = lambda name, use_cache=True: schema
    from embiggen.utils.normalize_kwargs import normalize_kwargs

    class Model:  # only named in error messages
        model_name = staticmethod(lambda: "Node2Vec SkipGram")
        library_name = staticmethod(lambda: "Ensmallen")
        task_name = staticmethod(lambda: "Node Embedding")

    records = []
    for key in KEYS:
        expected = schema[key]
        names = [expected] if isinstance(expected, str) else list(expected)
        for name in names:
            for value in CASES.get(name, []):
                try:
                    out = normalize_kwargs(Model, {key: value})[key]
                    records.append(dict(key=key, input=encode(value), output=encode(out)))
                except Exception as error:
                    records.append(dict(key=key, input=encode(value), error=type(error).__name__))
    try:
        normalize_kwargs(Model, {"no_such_kwarg": 1})
        unknown = None
    except Exception as error:
        unknown = type(error).__name__
    golden = dict(schema={key: schema[key] for key in KEYS}, records=records, unknown_kwarg_error=unknown)
    json.dump(golden, open(os.path.join(HERE, "normalize_kwargs_golden.json"), "w"), indent=1)
    print(f"wrote normalize_kwargs_golden.json: {len(records)} cases, unknown kwarg -> {unknown}")


if __name__ == "__main__":
    main()
