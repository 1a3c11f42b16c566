"""Measurement of the SURVEY 8(f) "next" rows on one B200 (not the headline bench: bench.py).

Workload: R-MAT 1 M nodes / 16 M edges (tables 2 x 512 MB, far above the 126 MB L2), L = 128,
w = 4, D = 100.  Kernel times are CUDA events on the stream the kernel is launched on; the
edge-prediction entry points take host buffers and are timed end to end (wall clock, copies
included).  One JSON object per line; `frac` = algorithmic bytes / measured HBM peak.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from embiggen_b200.edge_prediction import (DeviceFeatures, EdgeTransformerB200,  # noqa: E402
                                           PerceptronEdgePredictionB200, edge_metrics)
from embiggen_b200.engine import Engine  # noqa: E402
from embiggen_b200.graph_gpu import rmat_gpu  # noqa: E402

SEED, D, L, W, K = 42, 100, 128, 4, 10
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 7700.0


def emit(**record):
    print(json.dumps(record), flush=True)


def timed(stream, fn, repeat=3):
    """Median device time (ms) of fn() on `stream`, after one warm-up call."""
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(repeat):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


ONLY = os.environ.get("B2E_ROWS", "")  # e.g. B2E_ROWS=glove restricts the run (profiling)


def main():
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    graph = rmat_gpu(20, 16_000_000, n=1_000_000, seed=42, device=0)
    n, nnz = graph.get_number_of_nodes(), graph.indices.shape[0]
    rng = np.random.default_rng(0)
    weights = (rng.random(nnz) + 0.05).astype(np.float32)
    node_types = rng.integers(0, 4, n).astype(np.uint32)
    edge_types = rng.integers(0, 4, nnz).astype(np.uint32)
    chunk = 1 << 18
    emit(row="workload", graph="R-MAT scale 20, 1M nodes / 16M edges", n=n, nnz=int(nnz), walks_per_launch=chunk,
         walk_length=L, window=W, D=D, K=K, hbm_peak_gbs=PEAK)

    # ---- (f)-2: walk variants, steps/s of one launch of `chunk` walks ----
    variants = {
        "plain p=0.5 q=2": dict(),
        "weighted": dict(weights=True),
        "normalize_by_degree": dict(normalize_by_degree=True),
        "typed (node x3, edge x0.3)": dict(change_node_type_weight=3.0, change_edge_type_weight=0.3, types=True),
        "weighted + normalized + typed": dict(weights=True, normalize_by_degree=True, change_node_type_weight=3.0,
                                              change_edge_type_weight=0.3, types=True),
    }
    for name, opt in ({} if ONLY else variants).items():
        opt = dict(opt)
        use_weights, use_types = opt.pop("weights", False), opt.pop("types", False)
        with Engine("SkipGram", embedding_size=D, walk_length=L, window_size=W, iterations=1,
                    return_weight=2.0, explore_weight=0.5, chunk_walks=chunk, **opt) as engine:
            engine.load_csr(graph.indptr, graph.indices, weights if use_weights else None)
            if use_types:
                engine.load_types(node_types, edge_types)
            ws, ts = torch.cuda.Stream(device), torch.cuda.Stream(device)
            engine.set_streams(ws, ts)
            engine.reset_counters()
            calls = [0]

            def walk():
                engine.walk_chunk(SEED, calls[0] * chunk, chunk, 1, 0)
                calls[0] += 1
            ms = timed(ws, walk)
            c = engine.counters()
            steps = c["walk_steps"] / calls[0]
            trials = c["walk_trials"] / max(c["walk_steps"], 1)
            emit(row="f-2 walks", variant=name, ms_per_launch=ms, steps_per_s=steps / ms * 1e3,
                 trials_per_step=trials, searches_per_step=c["walk_searches"] / max(c["walk_steps"], 1))

    # ---- (f)-2 / (f)-3: SGD variants, pairs/s of one launch ----
    sgd = {
        "SkipGram pipelined (reference point)": dict(model="SkipGram"),
        "SkipGram stochastic_downsample_by_degree": dict(model="SkipGram", stochastic_downsample_by_degree=True),
        "Walklets SkipGram scale 2": dict(model="SkipGram", walklet_scale=2, window_size=1),
        "Walklets CBOW scale 3": dict(model="CBOW", walklet_scale=3, window_size=1),
    }
    for name, opt in ({} if ONLY else sgd).items():
        opt = dict(opt)
        model = opt.pop("model")
        kw = dict(embedding_size=D, walk_length=L, window_size=W, iterations=1, return_weight=2.0,
                  explore_weight=0.5, number_of_negative_samples=K, chunk_walks=chunk)
        kw.update(opt)
        with Engine(model, **kw) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            ws, ts = torch.cuda.Stream(device), torch.cuda.Stream(device)
            engine.set_streams(ws, ts)
            engine.init_tables(SEED)
            engine.walk_chunk(SEED, 0, chunk, 1, 0)
            engine.sync()
            engine.reset_counters()
            calls = [0]

            def train():
                engine.train_chunk(SEED, 0, 0.01)
                calls[0] += 1
            ms = timed(ts, train)
            c = engine.counters()
            pairs, targets = c["pairs"] / calls[0], c["targets"] / calls[0]
            bytes_ = targets * 2 * 4 * D  # rows read + written once per target
            emit(row="f-2/f-3 SGD", variant=name, ms_per_launch=ms, pairs_per_s=pairs / ms * 1e3,
                 algorithmic_gbs=bytes_ / ms / 1e6, frac=bytes_ / ms / 1e6 / PEAK)

    # ---- (f)-3: GloVe ----
    glove_walks = 1 << 17
    with Engine("GloVe", embedding_size=D, walk_length=L, window_size=W, iterations=1, return_weight=2.0,
                explore_weight=0.5, chunk_walks=glove_walks) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        ws, ts = torch.cuda.Stream(device), torch.cuda.Stream(device)
        engine.set_streams(ws, ts)
        engine.init_tables(SEED)
        t0 = time.perf_counter()
        triples = engine.cooccurrence(SEED, 0, glove_walks)
        cooc_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        triples = engine.cooccurrence(SEED, 0, glove_walks)
        cooc_s = min(cooc_s, time.perf_counter() - t0)
        slots = 2 * glove_walks * L * W
        emit(row="f-3 GloVe co-occurrence", walks=glove_walks, key_slots=slots, triples=triples, seconds=cooc_s,
             walk_pairs_per_s=slots / cooc_s)
        t0 = time.perf_counter()
        many = engine.cooccurrence(SEED, 0, 4 * glove_walks)  # four chunks: three sorted-list merges
        seconds = time.perf_counter() - t0
        emit(row="f-3 GloVe co-occurrence, 4 chunks merged", walks=4 * glove_walks, key_slots=4 * slots,
             triples=many, seconds=seconds, walk_pairs_per_s=4 * slots / seconds)
        triples = engine.cooccurrence(SEED, 0, glove_walks)
        ms = timed(ts, lambda: engine.glove_train(0.05))
        bytes_ = triples * 2 * 4 * D + triples * 12
        emit(row="f-3 GloVe SGD", triples=triples, ms_per_pass=ms, triples_per_s=triples / ms * 1e3,
             algorithmic_gbs=bytes_ / ms / 1e6, frac=bytes_ / ms / 1e6 / PEAK)

    if ONLY:
        return
    # ---- (f)-4: edge embeddings and the perceptron on resident features ----
    features = rng.normal(size=(n, D)).astype(np.float32)
    m = 4_000_000
    src, dst = rng.integers(0, n, m).astype(np.uint32), rng.integers(0, n, m).astype(np.uint32)
    with DeviceFeatures(features) as resident:
        for methods in ("Hadamard", "CosineSimilarity", ["Concatenate", "L2Distance"]):
            transformer = EdgeTransformerB200(methods)
            transformer.fit(resident)
            transformer.transform(src[:1000], dst[:1000])
            seconds = float("inf")
            for _ in range(2):  # wall clock of a host round trip: best of two
                t0 = time.perf_counter()
                out = transformer.transform(src, dst)
                seconds = min(seconds, time.perf_counter() - t0)
            emit(row="f-4 edge embedding (host edge list in, host matrix out)", methods=methods, edges=m,
                 width=out.shape[1], seconds=seconds, edges_per_s=m / seconds,
                 d2h_gbs=out.nbytes / seconds / 1e9)
        model = PerceptronEdgePredictionB200(edge_features=None, edge_embeddings="Hadamard", number_of_epochs=2,
                                             number_of_edges_per_mini_batch=4096)
        model.fit(graph, resident)  # first call: module load, allocator warm-up
        t0 = time.perf_counter()
        model.fit(graph, resident)
        seconds = time.perf_counter() - t0
        samples = 2 * (nnz // 4096) * 4096
        emit(row="f-4 perceptron fit (2 epochs, mini-batch 4096, Hadamard)", samples=samples, seconds=seconds,
             samples_per_s=samples / seconds, steps_per_s=2 * (nnz // 4096) / seconds, losses=model.get_losses())
        seconds = float("inf")
        for _ in range(3):
            t0 = time.perf_counter()
            scores = model.predict_proba(src, dst, resident)
            seconds = min(seconds, time.perf_counter() - t0)
        emit(row="f-4 perceptron predict (host edge list in, host scores out)", edges=m, seconds=seconds,
             edges_per_s=m / seconds, row_gather_gbs=m * 2 * 4 * D / seconds / 1e9,
             finite=bool(np.isfinite(scores).all()))


    # the reference's default perceptron: topological edge features only (no node features)
    t0 = time.perf_counter()
    metrics = edge_metrics(graph, src[:1_000_000], dst[:1_000_000], ["JaccardCoefficient", "AdamicAdar",
                                                                     "ResourceAllocationIndex"])
    seconds = time.perf_counter() - t0
    emit(row="f-4 edge features (Jaccard, Adamic-Adar, resource allocation; random pairs)", edges=1_000_000,
         seconds=seconds, edges_per_s=1_000_000 / seconds, finite=bool(np.isfinite(metrics).all()))
    model = PerceptronEdgePredictionB200(number_of_epochs=1, number_of_edges_per_mini_batch=4096)
    model.fit(graph)
    t0 = time.perf_counter()
    model.fit(graph)
    seconds = time.perf_counter() - t0
    samples = (nnz // 4096) * 4096
    emit(row="f-4 perceptron fit, default configuration (JaccardCoefficient, 1 epoch, mini-batch 4096)",
         samples=samples, seconds=seconds, samples_per_s=samples / seconds, losses=model.get_losses())


if __name__ == "__main__":
    main()
