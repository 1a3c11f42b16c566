import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from embiggen_b200.edge_prediction import DeviceFeatures, PerceptronEdgePredictionB200
from embiggen_b200.graph_gpu import rmat_gpu
graph = rmat_gpu(20, 16_000_000, n=1_000_000, seed=42, device=0)
rng = np.random.default_rng(0)
n = graph.get_number_of_nodes()
features = rng.normal(size=(n, 100)).astype(np.float32)
m = 4_000_000
src, dst = rng.integers(0, n, m).astype(np.uint32), rng.integers(0, n, m).astype(np.uint32)
with DeviceFeatures(features) as resident:
    model = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=1)
    model.fit(graph, resident)
    for i in range(4):
        t0 = time.perf_counter(); s = model.predict_proba(src, dst, resident); print("predict", i, time.perf_counter() - t0, flush=True)
    model.fit(graph, resident)
    for i in range(2):
        t0 = time.perf_counter(); s = model.predict_proba(src, dst, resident); print("predict after fit", i, time.perf_counter() - t0, flush=True)
