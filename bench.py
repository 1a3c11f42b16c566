#!/usr/bin/env python
"""Benchmark of the hot path: node2vec/DeepWalk walks + one SkipGram/CBOW SGD pass.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # CPU baseline (oracle port)

A *step* is one pass of the hot path over one batch: walks from every start node once
(n_src walks, one `iteration` of the reference's kwargs) followed by SGD over those walks.
Workload = BASELINE.json configs[1] (C2): DeepWalk SkipGram on Erdos-Renyi 1M nodes / 10M
edges, D=100, L=128, w=4, K=10.  `value` = context pairs/s (whole job, device-resident,
walk kernel of step k+1 overlapped with the SGD kernel of step k); walk steps/s is reported
beside it under "walk".  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: graph spec, model, kwargs
    "C2": dict(graph=("er", 1_000_000, 10_000_000), model="SkipGram", embedding_size=100,
               return_weight=1.0, explore_weight=1.0, iterations=10,
               label="DeepWalk SkipGram p=q=1, Erdos-Renyi 1M nodes / 10M edges"),
    "C3": dict(graph=("rmat", 24, 200_000_000, 10_000_000), model="SkipGram", embedding_size=100,
               return_weight=2.0, explore_weight=0.5, iterations=1,
               label="Node2Vec SkipGram p=0.5 q=2, R-MAT 10M nodes / 200M edges"),
    "C4": dict(graph=("rmat", 24, 200_000_000, 10_000_000), model="CBOW", embedding_size=128,
               return_weight=0.5, explore_weight=2.0, iterations=1,
               label="Node2Vec CBOW p=2 q=0.5, R-MAT 10M nodes / 200M edges, dim=128"),
    "C5": dict(graph=("rmat", 27, 2_000_000_000, 100_000_000), model="SkipGram", embedding_size=100,
               return_weight=2.0, explore_weight=0.5, iterations=1,
               label="Node2Vec SkipGram p=0.5 q=2, R-MAT 100M nodes / 2B edges"),
    # reduced shapes for quick functional runs; never the reported workload
    "small": dict(graph=("er", 100_000, 1_000_000), model="SkipGram", embedding_size=100,
                  return_weight=1.0, explore_weight=1.0, iterations=10,
                  label="DeepWalk SkipGram p=q=1, Erdos-Renyi 100k nodes / 1M edges (REDUCED)"),
    "small_n2v": dict(graph=("rmat", 17, 1_000_000, 100_000), model="SkipGram", embedding_size=100,
                      return_weight=2.0, explore_weight=0.5, iterations=1,
                      label="Node2Vec SkipGram p=0.5 q=2, R-MAT 100k nodes / 1M edges (REDUCED)"),
    "small_cbow": dict(graph=("rmat", 17, 1_000_000, 100_000), model="CBOW", embedding_size=128,
                       return_weight=0.5, explore_weight=2.0, iterations=1,
                       label="Node2Vec CBOW p=2 q=0.5, R-MAT 100k nodes / 1M edges, dim=128 (REDUCED)"),
}
COMMON = dict(walk_length=128, window_size=4, number_of_negative_samples=10, learning_rate=0.01,
              learning_rate_decay=0.9, clipping_value=6.0)
SEED = 42
GATHER_CEILING = 40.6e9  # measured random 4-byte gathers/s of a B200 over a 1.6 GB array
METRIC = "skipgram_context_pairs_per_s"


def load_graph(spec, device=None, world=1, local_rank=0, barrier=None):
    """Synthetic graph of the named shape (generator seed 42), cached as raw .npy files (tmpfs
    when available) and memory-mapped, so that the ranks of one node share one copy.

    With a GPU the graph is generated and its CSR built on the device (same graph as the numpy
    generators, tests/test_gpu_graph_build.py); the CPU reference arm falls back to numpy.
    With several ranks only local rank 0 generates; the others wait at `barrier` and map it."""
    from embiggen_b200.graph import CSRGraph, erdos_renyi, rmat
    tag = "_".join(str(x) for x in spec)
    default_cache = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    base = os.path.join(os.environ.get("B2E_CACHE", default_cache), f"b2e_graph_{tag}")
    paths = (base + ".indptr.npy", base + ".indices.npy")

    def cached():
        if all(os.path.exists(p) for p in paths):
            return CSRGraph(np.load(paths[0], mmap_mode="r"), np.load(paths[1], mmap_mode="r"), name=tag)
        return None

    graph = cached()
    if graph is None and local_rank == 0:
        if device is not None:
            from embiggen_b200.graph_gpu import erdos_renyi_gpu, rmat_gpu
            graph = erdos_renyi_gpu(spec[1], spec[2], seed=42, device=device) if spec[0] == "er" else \
                rmat_gpu(spec[1], spec[2], n=spec[3], seed=42, device=device)
        else:
            graph = erdos_renyi(spec[1], spec[2], seed=42) if spec[0] == "er" else \
                rmat(spec[1], spec[2], n=spec[3], seed=42)
        if world > 1 or graph.indices.shape[0] <= 1_000_000_000:  # one rank alone keeps 2 B edges in RAM
            try:
                for path, array in zip(paths, (graph.indptr, graph.indices)):
                    tmp = path + f".{os.getpid()}.tmp.npy"
                    np.save(tmp, array)
                    os.replace(tmp, path)
            except OSError:
                pass
    if barrier is not None:
        barrier()
    if graph is None:
        graph = cached()
    if graph is None:
        raise RuntimeError("the graph cache written by local rank 0 is not visible on this rank")
    return graph


def measured_traffic(config_name, walks_per_launch):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture of the same
    kernel and workload (profiles/traffic_<config>.json), scaled to this launch's walk count
    (the kernel streams walks, its traffic is linear in them).  None when no capture exists."""
    path = os.path.join(ROOT, "profiles", f"traffic_{config_name}.json")
    if not os.path.exists(path):
        return None, None
    record = json.load(open(path))
    per_walk = (record["dram_bytes_read"] + record["dram_bytes_write"]) / record["walks_per_launch"]
    return per_walk * walks_per_launch, record["source"]


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        clocks, max_clock, reasons = [], None, set()
        for line in self.samples:
            fields = [f.strip() for f in line.split(",")]
            if len(fields) < 6:
                continue
            try:
                clocks.append(float(fields[0]))
                max_clock = float(fields[1])
            except ValueError:
                continue
            for name, value in zip(self.NAMES, fields[2:6]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(clocks)) if clocks else None, "sm_max_mhz": max_clock,
                "reasons": sorted(reasons), "samples": len(clocks)}


def skipgram_bytes(counters, embedding_size, walk_length, centres):
    """Algorithmic bytes of the SGD kernel from the device counters.

    Every scored target row is read and written once (2 * 4D), every centre row likewise,
    every negative draw touches one 32 B alias sector, every centre reads its 4 B token.
    """
    row = 2 * 4 * embedding_size
    negatives_drawn = counters["pairs"] * COMMON["number_of_negative_samples"]
    return counters["targets"] * row + centres * (row + 4) + negatives_drawn * 32


def cbow_bytes(counters, embedding_size, centres):
    row = 2 * 4 * embedding_size
    negatives_drawn = centres * COMMON["number_of_negative_samples"]
    return (counters["pairs"] + counters["targets"]) * row + centres * 4 + negatives_drawn * 32


def cpu_baseline(graph, cfg, budget_s=15.0, threads=None):
    """Times the CPU oracle (multi-threaded Hogwild) on a bounded sample of the workload."""
    import oracle
    threads = threads or os.cpu_count() or 1
    oracle.set_threads(threads)
    n = graph.get_number_of_nodes()
    D = cfg["embedding_size"]
    thr, alias = oracle.alias_build(graph.indptr, 0.75)
    t0, t1 = oracle.init_tables(n, D, SEED)
    srcs = oracle.sources(graph.indptr)

    def run(first, count):
        begin = time.perf_counter()
        walks, wc = oracle.walks(graph.indptr, graph.indices, SEED, first, count,
                                 COMMON["walk_length"], cfg["return_weight"],
                                 cfg["explore_weight"], srcs=srcs)
        mid = time.perf_counter()
        stats = oracle.train(cfg["model"], walks, t0, t1, SEED, n, D, COMMON["window_size"],
                             COMMON["number_of_negative_samples"], COMMON["learning_rate"],
                             COMMON["clipping_value"], first_walk=first, thr=thr, alias=alias)
        end = time.perf_counter()
        return wc["steps"], stats["pairs"], mid - begin, end - mid

    _, pairs, tw, tt = run(0, 64 * threads)  # calibration (also warms the caches)
    rate = pairs / max(tw + tt, 1e-9)
    per_walk = pairs / (64 * threads)
    count = int(max(64 * threads, min(2_000_000, budget_s * rate / per_walk)))
    steps, pairs, tw, tt = run(64 * threads, count)
    oracle.set_threads(1)
    return {
        "value": pairs / (tw + tt), "unit": "pairs/s", "cores": threads, "kind": "port",
        "sample": f"{count} walks of the same workload ({pairs} pairs, {steps} walk steps): "
                  f"walks {tw:.2f} s + SGD {tt:.2f} s, OpenMP Hogwild over {threads} threads",
        "walk_steps_per_s": steps / max(tw, 1e-9), "sgd_pairs_per_s": pairs / max(tt, 1e-9),
    }


def run_reference(args, cfg):
    """--impl reference: the CPU implementation of the path on the host cores.

    The reference's own arithmetic lives in the `ensmallen` wheel, which is not vendored
    and not installable offline, so the timed implementation is the oracle port.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    graph = load_graph(cfg["graph"])
    threads = os.cpu_count() or 1
    oracle.set_threads(threads)
    n = graph.get_number_of_nodes()
    D = cfg["embedding_size"]
    thr, alias = oracle.alias_build(graph.indptr, 0.75)
    t0, t1 = oracle.init_tables(n, D, SEED)
    srcs = oracle.sources(graph.indptr)
    sample = max(64 * threads, args.reference_walks)
    times, pairs_total, steps_total = [], 0, 0
    for step in range(args.warmup + args.steps):
        begin = time.perf_counter()
        walks, wc = oracle.walks(graph.indptr, graph.indices, SEED, step * sample, sample,
                                 COMMON["walk_length"], cfg["return_weight"],
                                 cfg["explore_weight"], srcs=srcs)
        stats = oracle.train(cfg["model"], walks, t0, t1, SEED, n, D, COMMON["window_size"],
                             COMMON["number_of_negative_samples"], COMMON["learning_rate"],
                             COMMON["clipping_value"], first_walk=step * sample, thr=thr,
                             alias=alias)
        elapsed = time.perf_counter() - begin
        if step >= args.warmup:
            times.append(elapsed)
            pairs_total += stats["pairs"]
            steps_total += wc["steps"]
    total = sum(times)
    value = pairs_total / total
    sample_text = (f"each step = {sample} walks of the workload (walks + SGD), OpenMP Hogwild "
                   f"over {threads} threads")
    emit_result({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["label"], "name": args.config, **COMMON,
                   "embedding_size": D},
        "walk_steps_per_s": steps_total / total,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": sample_text},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    })


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from embiggen_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback.")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)

    graph = load_graph(cfg["graph"], device=local_rank, world=world, local_rank=local_rank,
                       barrier=dist.barrier if world > 1 else None)
    D = cfg["embedding_size"]
    L, w, K = COMMON["walk_length"], COMMON["window_size"], COMMON["number_of_negative_samples"]
    engine = Engine(cfg["model"], embedding_size=D, epochs=1, iterations=cfg["iterations"],
                    return_weight=cfg["return_weight"], explore_weight=cfg["explore_weight"],
                    chunk_walks=args.chunk_walks, device=local_rank, **COMMON)
    engine.load_csr(graph.indptr, graph.indices)
    n_src = engine.number_of_sources
    chunk = min(engine.chunk_capacity, n_src)  # walks per rank and step
    walk_stream, train_stream = torch.cuda.Stream(device), torch.cuda.Stream(device)
    engine.set_streams(walk_stream, train_stream)
    engine.init_tables(SEED)
    tables = engine.device_tables() if world > 1 else None
    lr = COMMON["learning_rate"]

    def first_id(step):  # weak scaling: every rank walks `chunk` ids of a world*chunk batch
        return step * chunk * world + rank

    def average_tables():
        engine.sync()
        for t in tables:
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize(device)

    def run_steps(first_step, count, events=None):
        """walk(k+1) on the walk stream overlaps train(k) on the train stream."""
        if count <= 0:
            return
        engine.walk_chunk(SEED, first_id(first_step), chunk, world, first_step & 1)
        for k in range(first_step, first_step + count):
            if k + 1 < first_step + count:
                engine.walk_chunk(SEED, first_id(k + 1), chunk, world, (k + 1) & 1)
            if events is not None:
                events[k - first_step][0].record(train_stream)
            engine.train_chunk(SEED, k & 1, lr)
            if events is not None:
                events[k - first_step][1].record(train_stream)
            if world > 1 and args.sync_interval and (k + 1) % args.sync_interval == 0:
                average_tables()

    def barrier():
        engine.sync()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    # ---- warm-up ----
    run_steps(0, args.warmup)
    barrier()

    # ---- timed region: exactly K steps (K walk launches + K SGD launches [+ all-reduces]) ----
    engine.reset_counters()
    launches_before = engine.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_launch = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.steps)]
    barrier()
    start.record(train_stream)
    walk_stream.wait_event(start)
    run_steps(args.warmup, args.steps, per_launch)
    stop.record(train_stream)
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    clocks = sampler.stop() if sampler else None
    counters = engine.counters()
    launches = engine.launch_count - launches_before
    sgd_ms = [a.elapsed_time(b) for a, b in per_launch]

    totals = torch.tensor([elapsed_ms, float(counters["pairs"]), float(counters["walk_steps"]),
                           float(launches)], dtype=torch.float64, device=device)
    if world > 1:
        max_ms = totals[:1].clone()
        dist.all_reduce(max_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(totals[1:], op=dist.ReduceOp.SUM)
        totals[0] = max_ms[0]
    elapsed_ms, pairs_all, walk_steps_all, launches_all = (float(x) for x in totals)
    value = pairs_all / (elapsed_ms * 1e-3)

    # ---- walk kernel alone (its own roofline line) ----
    walk_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(max(3, min(args.steps, 5)))]
    engine.reset_counters()
    for i, (a, b) in enumerate(walk_events):
        a.record(walk_stream)
        engine.walk_chunk(SEED, first_id(1000 + i), chunk, world, i & 1)
        b.record(walk_stream)
    barrier()
    walk_ms = [a.elapsed_time(b) for a, b in walk_events]
    walk_counters = engine.counters()

    peak, peak_source = measured_peak_gbs()
    centres = chunk * L  # every token of every walk is a centre once
    if cfg["model"] == "SkipGram":
        per_launch_bytes = skipgram_bytes(counters, D, L, centres * args.steps) / args.steps
    else:
        per_launch_bytes = cbow_bytes(counters, D, centres * args.steps) / args.steps
    sgd_avg_ms = float(np.mean(sgd_ms))
    achieved = per_launch_bytes / (sgd_avg_ms * 1e-3) / 1e9
    steps_per_launch = walk_counters["walk_steps"] / len(walk_events)
    trials = walk_counters["walk_trials"] / max(walk_counters["walk_steps"], 1)
    walk_bytes_per_step = 36 + 32 * max(trials, 1.0)  # + probe sectors (needs the oracle's count)
    walk_avg_ms = float(np.mean(walk_ms))
    walk_achieved = steps_per_launch * walk_bytes_per_step / (walk_avg_ms * 1e-3) / 1e9
    # A walk is a chain of dependent random gathers; when the CSR does not fit L2 the memory
    # system serves at most GATHER_CEILING random 4-byte gathers per second
    # (scripts/microbench_gather.cu, profiles/r01_microbench_random_gathers.txt).  Every step
    # needs at least its row bounds, its proposals and one probe per adjacency search.
    searches = walk_counters["walk_searches"] / max(walk_counters["walk_steps"], 1)
    gathers_per_step = 1.0 + max(trials, 1.0) + searches
    walk_gathers = steps_per_launch * gathers_per_step / (walk_avg_ms * 1e-3)

    kernel_name = ("skipgram_pipe_kernel" if cfg["model"] == "SkipGram" else "cbow_pipe_kernel") + \
        f"<{K + 1}>"
    traffic, traffic_source = measured_traffic(args.config, chunk)
    result = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["label"], "name": args.config, "embedding_size": D, **COMMON,
                   "walks_per_step_per_gpu": chunk, "pairs_per_step_per_gpu":
                   counters["pairs"] / args.steps, "l2": "inputs larger than L2 (tables "
                   f"{2 * graph.get_number_of_nodes() * D * 4 / 1e6:.0f} MB, random rows)",
                   "parallelism": f"dp{world}: start nodes sharded, CSR + tables replicated, "
                                  f"all-reduce AVG every {args.sync_interval} step(s)"
                   if world > 1 else "single GPU"},
        "walk_steps_per_s": walk_steps_all / (elapsed_ms * 1e-3),
        "gpu_launches": int(launches_all),
        "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_source,
                     "peak_source": peak_source, "avg_launch_ms": sgd_avg_ms,
                     "algorithmic_bytes_per_launch": per_launch_bytes},
        "walk": {"kernel": "walk_kernel", "steps_per_s_alone": steps_per_launch / (walk_avg_ms * 1e-3),
                 "avg_launch_ms": walk_avg_ms, "trials_per_step": trials,
                 "bytes_per_step": walk_bytes_per_step, "achieved_gbs": walk_achieved,
                 "frac": walk_achieved / peak,
                 "gathers_per_step_lower_bound": gathers_per_step,
                 "gathers_per_s_lower_bound": walk_gathers,
                 "gather_ceiling_per_s": GATHER_CEILING,
                 "gather_frac_lower_bound": walk_gathers / GATHER_CEILING},
        "clocks": clocks,
        "mean_pair_loss": counters["loss_sum"] / max(counters["pairs"], 1),
    }

    # ---- e2e: the reference-facing call with HOST buffers in and out, on every rank: CSR H2D +
    # init + K steps per GPU (+ the all-reduces) + tables D2H; max over ranks ----
    if not args.no_e2e:
        del engine, tables
        torch.cuda.synchronize(device)
        e2e_engine = Engine(cfg["model"], embedding_size=D, epochs=1,
                            iterations=args.steps * world, return_weight=cfg["return_weight"],
                            explore_weight=cfg["explore_weight"], chunk_walks=args.chunk_walks,
                            device=local_rank, **COMMON)
        n = graph.get_number_of_nodes()
        out0 = torch.empty((n, D), dtype=torch.float32, pin_memory=True).numpy()
        out1 = torch.empty((n, D), dtype=torch.float32, pin_memory=True).numpy()
        if world > 1:
            dist.barrier()
        begin = time.perf_counter()
        e2e_engine.load_csr(graph.indptr, graph.indices)
        if world > 1:
            t0, t1, _ = e2e_engine.fit_distributed(SEED, args.sync_interval)
            out0[:], out1[:] = t0, t1
        else:
            e2e_engine.fit(SEED, out0, out1)
        e2e_engine.sync()
        e2e_s = time.perf_counter() - begin
        stats = torch.tensor([e2e_s, float(e2e_engine.counters()["pairs"])], dtype=torch.float64,
                             device=device)
        if world > 1:
            slowest = stats[:1].clone()
            dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
            dist.all_reduce(stats[1:], op=dist.ReduceOp.SUM)
            stats[0] = slowest[0]
        e2e_s, e2e_pairs = float(stats[0]), float(stats[1])
        h2d = graph.indptr.nbytes + graph.indices.nbytes + 12 * n  # + sources, alias table
        result["e2e"] = {"value": e2e_pairs / e2e_s, "unit": "pairs/s",
                         "h2d_bytes_per_step": h2d / args.steps,
                         "d2h_bytes_per_step": (out0.nbytes + out1.nbytes) / args.steps,
                         "seconds": e2e_s,
                         "call": ("b2e_load_csr + b2e_fit" if world == 1 else
                                  "Engine.load_csr + Engine.fit_distributed on every rank") +
                                 f" (host buffers in/out, {args.steps} steps per GPU, epochs=1)"}
        e2e_engine.close()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(graph, cfg, budget_s=args.cpu_budget)
        emit_result(result)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Everything else a library may print there (NCCL's
    version banner ignores NCCL_DEBUG_FILE at NCCL_DEBUG=VERSION, torchrun children, CUDA
    warnings) is sent to stderr: file descriptor 1 is pointed at stderr for the whole run and
    the result is written to a saved duplicate of the original stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_result(record):
    line = (json.dumps(record) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    claim_stdout()
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    parser.add_argument("--chunk-walks", type=int, default=1 << 20)
    parser.add_argument("--sync-interval", type=int, default=1)
    parser.add_argument("--reference-walks", type=int, default=8192)
    parser.add_argument("--cpu-budget", type=float, default=15.0)
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
