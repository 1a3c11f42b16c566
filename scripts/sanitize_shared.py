"""compute-sanitizer driver for skipgram_shared_kernel (`shared_negatives=True`): the single-warp
and the production launch, compile-time and run-time K, rows that fill and do not fill a warp,
centre downsampling (window jumps), on a hub-heavy graph walked with return_weight 2 (repeated
tokens inside a window).  Prints 'sanitize shared ok' at the end."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from embiggen_b200.engine import Engine
from embiggen_b200.graph import rmat

hub = rmat(11, 20000, n=2000, seed=5)
for deterministic in (True, False):
    for extra in (dict(), dict(embedding_size=128, window_size=5), dict(number_of_negative_samples=7, window_size=7),
                  dict(stochastic_downsample_by_degree=True, embedding_size=5, number_of_negative_samples=0)):
        kw = dict(embedding_size=100, walk_length=40, window_size=4, iterations=1, epochs=1, return_weight=2.0,
                  explore_weight=0.5, number_of_negative_samples=10, shared_negatives=True,
                  deterministic=deterministic, chunk_walks=700)
        kw.update(extra)
        with Engine("SkipGram", **kw) as engine:
            engine.load_csr(hub.indptr, hub.indices)
            engine.init_tables(3)
            engine.walk_chunk(3, 0, 300 if deterministic else 700, 1, 0)
            engine.train_chunk(3, 0, 0.05)
            t0, t1 = engine.export_tables()
            assert np.isfinite(t0).all() and np.isfinite(t1).all() and engine.counters()["pairs"] > 0, kw
print("sanitize shared ok")
