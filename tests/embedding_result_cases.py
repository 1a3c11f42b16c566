"""`EmbeddingResult` (/root/reference/embiggen/utils/abstract_models/embedding_result.py:11-334)
beyond what the reference's own test file covers: what is validated (and when the scan is
skipped), what warns, what the single-embedding proxy exposes, the dump / load round trip.  One
table of outcomes for BOTH classes: the restatement (tests/test_embedder_api.py) and the
reference's (tests/real_embiggen_probe.py)."""
import warnings

import numpy as np
import pandas as pd


def run_cases(EmbeddingResult):
    def outcome(call):
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            try:
                value = call()
                text = type(value).__name__ if not isinstance(value, (bool, int, str, tuple, list)) else repr(value)
            except Exception as error:
                text = type(error).__name__
        if caught:
            text += " + " + caught[0].category.__name__
        return text

    ones = np.ones((4, 3), dtype=np.float32)
    frame = pd.DataFrame(ones * 2, index=list("abcd"))
    huge = np.zeros((1_000_001, 1), dtype=np.float32)
    huge[5, 0] = np.nan
    two = EmbeddingResult("M", node_embeddings=[ones, ones * 3])
    single = EmbeddingResult("M", node_embeddings=frame)
    return {
        "a bare array is wrapped in a list": outcome(lambda: len(EmbeddingResult("M", node_embeddings=ones).get_all_node_embedding())),
        "two node embeddings": outcome(lambda: (two.number_of_embeddings(), two.is_single_embedding())),
        "node and edge embeddings count together": outcome(
            lambda: EmbeddingResult("M", node_embeddings=ones, edge_embeddings=[ones, ones]).number_of_embeddings()),
        "index past the end": outcome(lambda: two.get_node_embedding_from_index(2)),
        "second embedding by index": outcome(lambda: float(two.get_node_embedding_from_index(1)[0, 0])),
        "asking for embeddings that were not given": outcome(lambda: two.get_all_edge_embedding()),
        "asking for a node type embedding that was not given": outcome(lambda: two.get_node_type_embedding_from_index(0)),
        "neither array nor frame": outcome(lambda: EmbeddingResult("M", node_embeddings=[[1.0, 2.0]])),
        "no rows": outcome(lambda: EmbeddingResult("M", node_embeddings=np.zeros((0, 3)))),
        "NaN": outcome(lambda: EmbeddingResult("M", node_embeddings=np.array([[1.0, np.nan]]))),
        "infinity": outcome(lambda: EmbeddingResult("M", edge_embeddings=np.array([[1.0, np.inf]]))),
        "NaN inside a frame": outcome(lambda: EmbeddingResult("M", node_embeddings=pd.DataFrame([[np.nan]]))),
        "all zeros warns": outcome(lambda: EmbeddingResult("M", node_embeddings=np.zeros((3, 2)))),
        "NaN beyond a million rows is not looked for": outcome(lambda: EmbeddingResult("M", node_embeddings=huge)),
        "no embedding at all": outcome(lambda: EmbeddingResult("M").number_of_embeddings()),
        "single embedding": outcome(lambda: (single.is_single_embedding(), type(single.get_single_embedding()).__name__)),
        "a single frame lends its methods": outcome(lambda: float(single.mean().iloc[0])),
        "a single frame lends to_numpy": outcome(lambda: single.to_numpy().shape),
        "several embeddings lend nothing": outcome(lambda: two.mean()),
        "method name": outcome(lambda: single.embedding_method_name),
        "dump keys": outcome(lambda: sorted(two.dump())),
        "dump and load": outcome(lambda: float(EmbeddingResult.load(two.dump()).get_node_embedding_from_index(1)[0, 0])),
    }


EXPECTED = {
    "a bare array is wrapped in a list": "1",
    "two node embeddings": "(2, False)",
    "node and edge embeddings count together": "3",
    "index past the end": "ValueError",
    "second embedding by index": "float",
    "asking for embeddings that were not given": "ValueError",
    "asking for a node type embedding that was not given": "ValueError",
    "neither array nor frame": "ValueError",
    "no rows": "ValueError",
    "NaN": "ValueError",
    "infinity": "ValueError",
    "NaN inside a frame": "ValueError",
    "all zeros warns": "EmbeddingResult + UserWarning",
    "NaN beyond a million rows is not looked for": "EmbeddingResult",
    "no embedding at all": "0",
    "single embedding": "(True, 'DataFrame')",
    "a single frame lends its methods": "float",
    "a single frame lends to_numpy": "(4, 3)",
    "several embeddings lend nothing": "AttributeError",
    "method name": "'M'",
    "dump keys": "['edge_embeddings', 'edge_type_embeddings', 'embedding_method_name', 'node_embeddings', 'node_type_embeddings']",
    "dump and load": "float",
}
