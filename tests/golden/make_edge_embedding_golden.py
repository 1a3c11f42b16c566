"""Generates tests/golden/edge_embedding_golden.npz from the REFERENCE's own functions.

/root/reference/embiggen/embedding_transformers/edge_transformer.py cannot be imported here (it
imports `ensmallen` and `userinput` at module level), but its twelve edge-embedding methods are
pure numpy.  This script parses that file with `ast`, executes only the module-level
`get_*` function definitions in a namespace that holds numpy, and evaluates them on seeded
inputs.  Nothing of the reference is copied into the repository: the functions run from where
they lie, and only their outputs are committed (this script is the generating script the
contract asks for).  Run it in the build container (the GPU box has no /root/reference).
"""
import ast
import os
import sys

import numpy as np

REFERENCE = "/root/reference/embiggen/embedding_transformers/edge_transformer.py"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "edge_embedding_golden.npz")

# EdgeTransformer.methods (edge_transformer.py:337-350): name -> function
METHODS = {
    "Hadamard": "get_hadamard_edge_embedding", "Sum": "get_sum_edge_embedding",
    "Average": "get_average_edge_embedding", "L1": "get_l1_edge_embedding",
    "AbsoluteL1": "get_absolute_l1_edge_embedding", "SquaredL2": "get_squared_l2_edge_embedding",
    "L2": "get_l2_edge_embedding", "Concatenate": "get_concatenate_edge_embedding",
    "Min": "get_min_edge_embedding", "Max": "get_max_edge_embedding",
    "L2Distance": "get_l2_distance", "CosineSimilarity": "get_cosine_similarity",
}


def reference_functions():
    tree = ast.parse(open(REFERENCE).read(), REFERENCE)
    functions = [node for node in tree.body
                 if isinstance(node, ast.FunctionDef) and node.name.startswith("get_")]
    namespace = {"np": np}
    exec(compile(ast.Module(body=functions, type_ignores=[]), REFERENCE, "exec"), namespace)
    return {name: namespace[function] for name, function in METHODS.items()}


def main():
    functions = reference_functions()
    rng = np.random.default_rng(20261017)
    out = {}
    for case, (n, dim, m) in enumerate([(50, 8, 100), (120, 100, 48), (40, 5, 64), (64, 128, 24)]):
        features = rng.normal(size=(n, dim)).astype(np.float32)
        features[3] = 0.0            # a zero row: CosineSimilarity's norm clamp (1e-6)
        features[5] = features[4]    # identical rows: zero distances
        src = rng.integers(0, n, m).astype(np.uint32)
        dst = rng.integers(0, n, m).astype(np.uint32)
        src[:4], dst[:4] = [3, 4, 3, 7], [7, 5, 3, 7]
        out[f"case{case}_features"] = features
        out[f"case{case}_src"] = src
        out[f"case{case}_dst"] = dst
        for name, function in functions.items():
            out[f"case{case}_{name}"] = np.asarray(function(features[src], features[dst]), dtype=np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), {k: v.shape for k, v in out.items() if k.startswith("case1_")})


if __name__ == "__main__":
    if not os.path.exists(REFERENCE):
        sys.exit("the reference tree is not available here")
    main()
