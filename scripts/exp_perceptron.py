"""Where does a perceptron step go?  Steps/s against the mini-batch size (launch-bound if flat)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from embiggen_b200.edge_prediction import DeviceFeatures, PerceptronEdgePredictionB200  # noqa: E402
from embiggen_b200.graph_gpu import rmat_gpu  # noqa: E402

if os.environ.get("GRAPH") == "big":
    graph = rmat_gpu(20, 16_000_000, n=1_000_000, seed=42, device=0)
else:
    graph = rmat_gpu(18, 2_000_000, n=200_000, seed=42, device=0)
nnz = graph.indices.shape[0]
features = np.random.default_rng(0).normal(size=(graph.get_number_of_nodes(), 100)).astype(np.float32)
with DeviceFeatures(features) as resident:
    for batch in (int(b) for b in os.environ.get("BATCHES", "1024,4096,16384,65536").split(",")):
        model = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=1,
                                             number_of_edges_per_mini_batch=batch)
        model.fit(graph, resident)
        t0 = time.perf_counter()
        model.fit(graph, resident)
        seconds = time.perf_counter() - t0
        steps = max(1, nnz // batch)
        print(f"batch {batch:6d}: {steps:5d} steps in {seconds:.4f} s = {seconds / steps * 1e6:8.1f} us/step, "
              f"{steps * batch / seconds / 1e6:8.1f} M samples/s", flush=True)
