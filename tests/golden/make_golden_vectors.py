"""Regenerates tests/golden/hot_path_vectors.npz: outputs of the CPU oracle (the normative
definition of the path) on BASELINE config C1's graph.

    python tests/golden/make_golden_vectors.py

The reference holds no golden vectors for this path (parity unpinned against
Ensmallen), so these fixtures pin the *specification*: any change to the Philox layout, the
rejection sampler, the alias construction or the update order shows up as a diff here, for the
oracle (tests/test_golden.py, CPU) and for the CUDA path through the C ABI (-m gpu).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = dict(seed=42, walk_length=32, n_walks=64, first_walk=1000, return_weight=0.25,
             explore_weight=4.0, embedding_size=20, window_size=3, negatives=5,
             learning_rate=0.05, train_walks=300)


def digest(array):
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).hexdigest()


def build(graph_path):
    data = np.load(graph_path)
    indptr, indices = data["indptr"], data["indices"]
    c = CASES
    n = indptr.shape[0] - 1
    out = {}
    for name, (rw, ew) in dict(node2vec=(c["return_weight"], c["explore_weight"]),
                               deepwalk=(1.0, 1.0)).items():
        walks, counters = oracle.walks(indptr, indices, c["seed"], c["first_walk"], c["n_walks"],
                                       c["walk_length"], rw, ew)
        out[f"walks_{name}"] = walks
        out[f"walk_counters_{name}"] = np.array(
            [counters[k] for k in ("steps", "trials", "first_order", "searches")], dtype=np.uint64)
    thr, alias = oracle.alias_build(indptr, 0.75)
    out["alias_head"] = np.stack([thr[:64], alias[:64]])
    out["alias_digest"] = np.array([digest(thr), digest(alias)])
    t0, t1 = oracle.init_tables(n, c["embedding_size"], c["seed"])
    out["init_digest"] = np.array([digest(t0[:, :c["embedding_size"]]), digest(t1[:, :c["embedding_size"]])])
    out["init_head"] = t0[:4, :c["embedding_size"]].copy()
    train_walks, _ = oracle.walks(indptr, indices, c["seed"], 0, c["train_walks"], c["walk_length"],
                                  c["return_weight"], c["explore_weight"])
    for model in ("SkipGram", "CBOW"):
        a, b = oracle.init_tables(n, c["embedding_size"], c["seed"])
        stats = oracle.train(model, train_walks, a, b, c["seed"], n, c["embedding_size"],
                             c["window_size"], c["negatives"], c["learning_rate"], thr=thr, alias=alias)
        D = c["embedding_size"]
        out[f"tables_digest_{model}"] = np.array([digest(a[:, :D]), digest(b[:, :D])])
        out[f"tables_head_{model}"] = np.stack([a[:8, :D], b[:8, :D]])
        out[f"stats_{model}"] = np.array([stats["pairs"], stats["targets"]], dtype=np.uint64)
        out[f"loss_{model}"] = np.array([stats["loss_sum"]])
    return out


if __name__ == "__main__":
    golden = os.path.join(ROOT, "tests", "golden")
    vectors = build(os.path.join(golden, "small_ppi_csr.npz"))
    np.savez_compressed(os.path.join(golden, "hot_path_vectors.npz"), **vectors)
    print("wrote hot_path_vectors.npz:", ", ".join(sorted(vectors)))
