#!/usr/bin/env python
"""Benchmark of the hot path: node2vec/DeepWalk walks + SkipGram/CBOW SGD over them.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the path on the host cores

Workload (default): BASELINE.json's headline shape, config C5 -- Node2Vec SkipGram p=0.5 q=2 on
a synthetic R-MAT graph of 100 M nodes / 2 B edges, D=100, L=128, w=4, K=10 -- which fits one
B200 (125 GB of 180).  When the host has too little memory for it the run falls back to C3 (the
same model on R-MAT 10 M / 200 M) and says so in `config.workload`.

A *step* is one pass of the hot path over one batch: `chunk` walks per GPU (2^20) from
consecutive start nodes, then SGD over those walks; the walk kernel of step k+1 overlaps the SGD
kernel of step k on a second stream.  `value` = context pairs/s of the whole job, device
resident (CUDA events, max over ranks); with N > 1 GPUs the replicas are averaged every
`--sync-interval` steps by the peer-memory exchange kernel, inside the timed region.  `e2e` = the
named job in full through the reference-facing call -- host CSR in, one epoch over every start
node, host tables out.  One JSON line on stdout (rank 0).
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
    # name: graph spec, model, kwargs (BASELINE.json configs[1..4])
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
HEADLINE, FALLBACK = "C5", "C3"
HEADLINE_HOST_GB = 150  # C5 on the host: 17 GB of CSR (+ its tmpfs copy) and 80 GB of tables
GATHER_CEILING = 40.6e9  # measured random 4-byte gathers/s of a B200 over a 1.6 GB array
METRIC = "skipgram_context_pairs_per_s"


def host_memory_gb():
    try:
        with open("/proc/meminfo") as handle:
            for line in handle:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def choose_config(requested):
    """(name, note): the headline shape unless the host cannot hold it (both arms decide alike)."""
    if requested != "auto":
        return requested, None
    available = host_memory_gb()
    if available >= HEADLINE_HOST_GB or os.environ.get("B2E_FORCE_HEADLINE"):
        return HEADLINE, None
    # (no measured number in the text: both arms must print the same config)
    return FALLBACK, (f"fallback from {HEADLINE}: this host has less than {HEADLINE_HOST_GB} GB of memory available, "
                      "which the 100M-node tables need on the CPU arm")


# engine options outside the named workload (--shared-negatives); empty by default, so that the
# `config` of the default run is the one the reference arm prints
OPTIONS = {}


def config_dict(name, cfg, note):
    """The `config` object: the same in both arms (the driver compares them)."""
    n = cfg["graph"][3] if cfg["graph"][0] == "rmat" else cfg["graph"][1]
    tables_mb = 2 * n * cfg["embedding_size"] * 4 / 1e6
    return {"workload": cfg["label"] + (f" [{note}]" if note else ""), "name": name,
            "objective": cfg["model"], "embedding_size": cfg["embedding_size"],
            "return_weight": cfg["return_weight"], "explore_weight": cfg["explore_weight"], **COMMON, **OPTIONS,
            "l2": f"inputs larger than L2 (tables {tables_mb:.0f} MB, rows gathered at random; no flush)"}


def cache_paths(spec):
    tag = "_".join(str(x) for x in spec)
    default_cache = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    base = os.path.join(os.environ.get("B2E_CACHE", default_cache), f"b2e_graph_{tag}")
    return tag, (base + ".indptr.npy", base + ".indices.npy")


def load_graph(spec, device=None, local_rank=0, barrier=None):
    """Synthetic graph of the named shape (generator seed 42), cached as raw .npy files (tmpfs
    when available) and memory-mapped, so that the ranks of one node -- and the two arms of one
    driver run -- share one copy.

    Ours (`device` given): generated and built on the GPU (csrc/graph_build.cu).  The reference
    arm (`device` None): generated on the host cores by oracle/graphgen.c; it never loads the
    product library.  All generators implement one definition and tests compare them.  With several
    ranks only local rank 0 generates; the others wait at `barrier` and map the files."""
    from embiggen_b200.graph import CSRGraph
    tag, paths = cache_paths(spec)

    def cached():
        if all(os.path.exists(p) for p in paths):
            return CSRGraph(np.load(paths[0], mmap_mode="r"), np.load(paths[1], mmap_mode="r"), name=tag)
        return None

    graph = cached()
    seconds = 0.0
    if graph is None and local_rank == 0:
        begin = time.perf_counter()
        n = spec[3] if spec[0] == "rmat" else spec[1]
        if device is not None:
            from embiggen_b200.graph_gpu import erdos_renyi_gpu, rmat_gpu
            built = erdos_renyi_gpu(spec[1], spec[2], seed=42, device=device) if spec[0] == "er" else \
                rmat_gpu(spec[1], spec[2], n=spec[3], seed=42, device=device)
            arrays = (built.indptr, built.indices)
            del built
        else:
            import oracle
            oracle.set_threads(os.cpu_count() or 1)
            arrays = oracle.synthetic_csr("er", n, spec[2], seed=42) if spec[0] == "er" else \
                oracle.synthetic_csr("rmat", n, spec[2], scale=spec[1], seed=42)
        seconds = time.perf_counter() - begin
        try:
            for path, array in zip(paths, arrays):
                tmp = path + f".{os.getpid()}.tmp.npy"
                np.save(tmp, array)
                os.replace(tmp, path)
            del arrays
        except OSError:
            graph = CSRGraph(arrays[0], arrays[1], name=tag)
    if barrier is not None:
        barrier()
    if graph is None:
        graph = cached()
    if graph is None:
        raise RuntimeError("the graph cache written by local rank 0 is not visible on this rank")
    return graph, seconds


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


def skipgram_bytes(counters, embedding_size, centres):
    """Algorithmic bytes of the SGD kernel from the device counters.

    Every scored target row is read and written once (2 * 4D), every centre row likewise,
    every negative draw touches one 32 B alias sector, every centre reads its 4 B token.
    """
    row = 2 * 4 * embedding_size
    negatives_drawn = counters["pairs"] * COMMON["number_of_negative_samples"]
    return counters["targets"] * row + centres * (row + 4) + negatives_drawn * 32


def shared_skipgram_bytes(counters, embedding_size, centres):
    """--shared-negatives: per centre its own T0 row and the one T1 row that enters (and, updated,
    leaves) the window are read and written once, as is every valid negative's row."""
    row = 2 * 4 * embedding_size
    valid_negatives = counters["targets"] - counters["pairs"]
    return valid_negatives * row + centres * (2 * row + 4) + centres * COMMON["number_of_negative_samples"] * 32


def cbow_bytes(counters, embedding_size, centres):
    row = 2 * 4 * embedding_size
    negatives_drawn = centres * COMMON["number_of_negative_samples"]
    return (counters["pairs"] + counters["targets"]) * row + centres * 4 + negatives_drawn * 32


class PinnedTables:
    """Two (n, D) float32 host tables, page-locked with cudaHostRegister for the duration of the
    e2e call and handed back to the OS afterwards (torch's pinned allocator would keep 80 GB of
    C5 cached while the CPU baseline needs the same again)."""

    def __init__(self, n, D):
        from embiggen_b200 import _lib
        self.lib = _lib.load()
        self.tables = [np.empty((n, D), dtype=np.float32) for _ in range(2)]
        self.registered = []
        for table in self.tables:
            table.fill(0.0)  # touch every page before it is locked
            if self.lib.b2e_host_register(table.ctypes.data, table.nbytes) == 0:
                self.registered.append(table)

    def release(self):
        for table in self.registered:
            self.lib.b2e_host_unregister(table.ctypes.data)
        self.registered, self.tables = [], []


class CpuPath:
    """The path on the host cores: oracle/ (plain C + OpenMP Hogwild, `fast_math` arithmetic: a
    vectorised dot and libm exp instead of the warp-shaped bit-exact forms the parity tests use)."""

    def __init__(self, graph, cfg, threads=None):
        import oracle
        self.oracle, self.graph, self.cfg = oracle, graph, cfg
        self.threads = threads or os.cpu_count() or 1
        oracle.set_threads(self.threads)
        self.n = graph.get_number_of_nodes()
        self.D = cfg["embedding_size"]
        self.thr, self.alias = oracle.alias_build(graph.indptr, 0.75)
        self.t0, self.t1 = oracle.init_tables(self.n, self.D, SEED)
        self.srcs = oracle.sources(graph.indptr)

    def step(self, first, count):
        """walks + SGD over `count` walks; (walk steps, pairs, walk seconds, SGD seconds)"""
        o, cfg = self.oracle, self.cfg
        begin = time.perf_counter()
        walks, wc = o.walks(self.graph.indptr, self.graph.indices, SEED, first, count, COMMON["walk_length"],
                            cfg["return_weight"], cfg["explore_weight"], srcs=self.srcs, undirected=True)
        mid = time.perf_counter()
        stats = o.train(cfg["model"], walks, self.t0, self.t1, SEED, self.n, self.D, COMMON["window_size"],
                        COMMON["number_of_negative_samples"], COMMON["learning_rate"],
                        COMMON["clipping_value"], first_walk=first, thr=self.thr, alias=self.alias,
                        fast_math=True, shared_negatives=bool(OPTIONS.get("shared_negatives")))
        end = time.perf_counter()
        return wc["steps"], stats["pairs"], mid - begin, end - mid

    def close(self):
        self.oracle.set_threads(1)


def cpu_baseline(graph, cfg, budget_s=15.0):
    """Times the CPU path on a bounded sample of the workload (rank 0, N = 1 only)."""
    path = CpuPath(graph, cfg)
    threads = path.threads
    _, pairs, tw, tt = path.step(0, 64 * threads)  # calibration (also warms the caches)
    rate = pairs / max(tw + tt, 1e-9)
    per_walk = pairs / (64 * threads)
    count = int(max(64 * threads, min(2_000_000, budget_s * rate / per_walk)))
    steps, pairs, tw, tt = path.step(64 * threads, count)
    path.close()
    return {
        "value": pairs / (tw + tt), "unit": "pairs/s", "cores": threads, "kind": "port",
        "sample": f"{count} walks of the same workload ({pairs} pairs, {steps} walk steps): "
                  f"walks {tw:.2f} s + SGD {tt:.2f} s, OpenMP Hogwild over {threads} threads, "
                  "vectorised float32 arithmetic (oracle fast_math)",
        "walk_steps_per_s": steps / max(tw, 1e-9), "sgd_pairs_per_s": pairs / max(tt, 1e-9),
    }


def ensmallen_probe():
    """The reference's own engine, when the wheel exists on this box (it never has so far)."""
    for extra in (os.path.join(ROOT, "baseline", "_ref"),):
        if os.path.isdir(extra) and extra not in sys.path:
            sys.path.insert(0, extra)
    try:
        import ensmallen  # noqa: F401
        return True, getattr(ensmallen, "__version__", "unknown")
    except Exception as error:  # ModuleNotFoundError here and on the GPU boxes (profiles/r02a_gpu_box_probe.txt)
        return False, f"{type(error).__name__}: {error}"


def run_ensmallen(args, name, cfg, note):
    """`Node2VecSkipGramEnsmallen(...).fit_transform` on the same graph, timed wall-clock on all host
    cores.  Only reachable where `import ensmallen` works; whole-job pairs/s."""
    import ensmallen
    from embiggen.embedders.ensmallen_embedders import Node2VecCBOWEnsmallen, Node2VecSkipGramEnsmallen
    graph, _ = load_graph(cfg["graph"])
    n = graph.get_number_of_nodes()
    rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(graph.indptr))
    keep = rows < graph.indices
    path = os.path.join(os.environ.get("B2E_CACHE", "/tmp"), f"b2e_edges_{name}.tsv")
    np.savetxt(path, np.stack([rows[keep], np.asarray(graph.indices)[keep]], axis=1), fmt="%d", delimiter="\t")
    g = ensmallen.Graph.from_csv(edge_path=path, directed=False, edge_list_header=False, sources_column_number=0,
                                 destinations_column_number=1, edge_list_numeric_node_ids=True,
                                 number_of_nodes=n, name=name)
    model_class = Node2VecSkipGramEnsmallen if cfg["model"] == "SkipGram" else Node2VecCBOWEnsmallen
    kwargs = dict(embedding_size=cfg["embedding_size"], epochs=1, iterations=cfg["iterations"],
                  return_weight=cfg["return_weight"], explore_weight=cfg["explore_weight"], max_neighbours=None,
                  verbose=False, **COMMON)
    from embiggen_b200.engine import pairs_per_walk
    walks = cfg["iterations"] * int((np.diff(graph.indptr) > 0).sum())
    pairs = walks * pairs_per_walk(COMMON["walk_length"], COMMON["window_size"])
    times = []
    for _ in range(max(1, min(args.steps, 3))):
        begin = time.perf_counter()
        model_class(**kwargs).fit_transform(g, return_dataframe=False)
        times.append(time.perf_counter() - begin)
    value = pairs / float(np.median(times))
    threads = os.cpu_count() or 1
    sample = f"one epoch over every start node ({walks} walks), wall clock, median of {len(times)}, rayon over {threads} threads"
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.median(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(name, cfg, note),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "ensmallen", "sample": sample},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args, name, cfg, note):
    """--impl reference: the path on the host cores, all of them.

    The reference's own arithmetic lives in the `ensmallen` wheel, which is neither vendored nor
    installable offline (probe logged under profiles/); when it is importable it is what gets
    timed, otherwise the oracle port.  Nothing here loads libb2e.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    have_wheel, wheel_note = ensmallen_probe()
    if have_wheel and not os.environ.get("B2E_REFERENCE_PORT"):
        try:
            emit_result(run_ensmallen(args, name, cfg, note))
            return
        except Exception as error:
            wheel_note = f"ensmallen importable but the run failed ({type(error).__name__}: {error})"
    graph, generation_s = load_graph(cfg["graph"])
    path = CpuPath(graph, cfg)
    threads = path.threads
    calibration = 64 * threads
    _, pairs, tw, tt = path.step(0, calibration)
    per_step_s = args.reference_step_seconds
    sample = int(np.clip(per_step_s * (pairs / max(tw + tt, 1e-9)) / (pairs / calibration), 1024, 1 << 18))
    if args.reference_walks:
        sample = args.reference_walks
    times, pairs_total, steps_total = [], 0, 0
    for step in range(args.warmup + args.steps):
        steps, pairs, tw, tt = path.step(calibration + step * sample, sample)
        if step >= args.warmup:
            times.append(tw + tt)
            pairs_total += pairs
            steps_total += steps
    path.close()
    total = sum(times)
    value = pairs_total / total
    sample_text = (f"each step = {sample} walks of the workload (walks + SGD), OpenMP Hogwild over "
                   f"{threads} threads, vectorised float32 arithmetic (oracle fast_math); "
                   f"graph generated on the host in {generation_s:.0f} s (0 = cached); ensmallen: {wheel_note}")
    emit_result({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(len(times), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, note),
        "walk_steps_per_s": steps_total / total,
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": sample_text},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    })


def run_ours(args, name, cfg, note):
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

    graph, generation_s = load_graph(cfg["graph"], device=local_rank, local_rank=local_rank,
                                     barrier=dist.barrier if world > 1 else None)
    n = graph.get_number_of_nodes()
    D = cfg["embedding_size"]
    L, w, K = COMMON["walk_length"], COMMON["window_size"], COMMON["number_of_negative_samples"]
    engine = Engine(cfg["model"], embedding_size=D, epochs=1, iterations=cfg["iterations"],
                    return_weight=cfg["return_weight"], explore_weight=cfg["explore_weight"],
                    chunk_walks=args.chunk_walks, device=local_rank, **COMMON, **OPTIONS)
    begin = time.perf_counter()
    engine.load_csr(graph.indptr, graph.indices)
    load_s = time.perf_counter() - begin
    n_src = engine.number_of_sources
    chunk = min(engine.chunk_capacity, n_src)  # walks per rank and step
    walk_stream, train_stream = torch.cuda.Stream(device), torch.cuda.Stream(device)
    engine.set_streams(walk_stream, train_stream)
    engine.init_tables(SEED)
    if world > 1:
        engine.open_exchange()
    lr = COMMON["learning_rate"]
    exchange_s = []

    def first_id(step):  # weak scaling: every rank walks `chunk` ids of a world*chunk batch
        return step * chunk * world + rank

    def average_tables():
        """Engine.average() spelled out, with CUDA events around the exchange kernel and the host
        clock around the whole step (barriers included, the wait for this rank's SGD excluded)."""
        engine.sync()
        begin = time.perf_counter()
        dist.barrier()
        events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        events[0].record(train_stream)
        engine.exchange_average()
        events[1].record(train_stream)
        engine.sync()
        dist.barrier()
        exchange_s.append((time.perf_counter() - begin, events))

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
            if world > 1 and args.sync_interval and (k + 1 - first_step) % args.sync_interval == 0:
                average_tables()

    def barrier():
        engine.sync()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    # ---- warm-up ----
    run_steps(0, args.warmup)
    barrier()

    # ---- timed region: exactly K steps (K walk launches + K SGD launches [+ exchange kernels]) ----
    engine.reset_counters()
    exchange_s.clear()
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
    timed_exchanges = list(exchange_s)

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
        per_launch_bytes = (shared_skipgram_bytes if OPTIONS.get("shared_negatives") else skipgram_bytes)(
            counters, D, centres * args.steps) / args.steps
    else:
        per_launch_bytes = cbow_bytes(counters, D, centres * args.steps) / args.steps
    sgd_avg_ms = float(np.mean(sgd_ms))
    achieved = per_launch_bytes / (sgd_avg_ms * 1e-3) / 1e9
    steps_per_launch = walk_counters["walk_steps"] / len(walk_events)
    per_step = lambda key: walk_counters[key] / max(walk_counters["walk_steps"], 1)  # noqa: E731
    second_order = not (cfg["return_weight"] == 1.0 and cfg["explore_weight"] == 1.0)
    trials, searches, probes = per_step("walk_trials"), per_step("walk_searches"), per_step("walk_probes")
    # 32 B for the row bounds + 4 B of output per step, one 32 B sector per proposal
    # and per gather of an adjacency check (filter word, row bounds, bisection step: counted on the
    # device); DeepWalk has exactly one proposal per step and no checks
    walk_bytes_per_step = 36 + 32 * (trials if second_order else 1.0) + 32 * probes
    walk_avg_ms = float(np.mean(walk_ms))
    walk_achieved = steps_per_launch * walk_bytes_per_step / (walk_avg_ms * 1e-3) / 1e9
    # A walk is a chain of dependent random gathers; when the CSR does not fit L2 the memory
    # system serves at most GATHER_CEILING random 4-byte gathers per second
    # (scripts/microbench_gather.cu, profiles/r01_microbench_random_gathers.txt).
    gathers_per_step = 1.0 + (trials if second_order else 1.0) + probes
    walk_gathers = steps_per_launch * gathers_per_step / (walk_avg_ms * 1e-3)

    kernel_name = ("skipgram_pipe_kernel" if cfg["model"] == "SkipGram" else "cbow_pipe_kernel") + \
        f"<{K + 1}>"
    traffic, traffic_source = measured_traffic(name, chunk)
    if OPTIONS.get("shared_negatives"):
        kernel_name = f"skipgram_shared_kernel<{K}>"
        traffic, traffic_source = measured_traffic(name + "_shared", chunk)
    result = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(name, cfg, note),
        "run": {"walks_per_step_per_gpu": chunk, "pairs_per_step_per_gpu": counters["pairs"] / args.steps,
                "start_nodes": n_src, "graph_generation_s": generation_s, "load_csr_s": load_s,
                "parallelism": (f"dp{world}: start nodes sharded, CSR + tables replicated, replicas averaged "
                                f"by the peer-memory exchange kernel every {args.sync_interval} step(s)")
                if world > 1 else "single GPU"},
        "walk_steps_per_s": walk_steps_all / (elapsed_ms * 1e-3),
        "gpu_launches": int(launches_all),
        "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_source,
                     "peak_source": peak_source, "avg_launch_ms": sgd_avg_ms,
                     "algorithmic_bytes_per_launch": per_launch_bytes},
        "walk": {"kernel": "walk_kernel", "steps_per_s_alone": steps_per_launch / (walk_avg_ms * 1e-3),
                 "steps_per_s_alone_all_gpus": world * steps_per_launch / (walk_avg_ms * 1e-3),  # rank 0's rate x N
                 "avg_launch_ms": walk_avg_ms, "trials_per_step": trials, "searches_per_step": searches,
                 "probe_gathers_per_step": probes,
                 "filter_rejects_per_search": walk_counters["walk_filter_rejects"] / max(walk_counters["walk_searches"], 1),
                 "bytes_per_step": walk_bytes_per_step, "achieved_gbs": walk_achieved,
                 "frac": walk_achieved / peak,
                 "gathers_per_step": gathers_per_step, "gathers_per_s": walk_gathers,
                 "gather_ceiling_per_s": GATHER_CEILING, "gather_frac": walk_gathers / GATHER_CEILING},
        "clocks": clocks,
        "mean_pair_loss": counters["loss_sum"] / max(counters["pairs"], 1),
    }
    if world > 1:
        live = 2 * n * D * 4
        kernel_ms = [a.elapsed_time(b) for _, (a, b) in timed_exchanges]
        mean_ms = float(np.mean(kernel_ms)) if kernel_ms else None
        result["exchange"] = {
            "kernel": f"exchange_average_kernel<{world}>", "sync_interval_steps": args.sync_interval,
            "syncs_in_timed_region": len(timed_exchanges), "kernel_ms_per_sync": mean_ms,
            "wall_ms_per_sync": 1e3 * float(np.mean([w for w, _ in timed_exchanges])) if timed_exchanges else None,
            "nvlink_bytes_per_direction_per_gpu_per_sync": 2 * live * (world - 1) / world,
            "nvlink_gbs_per_direction_per_gpu": None if not mean_ms else 2 * live * (world - 1) / world / (mean_ms * 1e-3) / 1e9,
            "note": "rank 0; kernel time by CUDA events on the train stream, wall clock = barrier + kernel + "
                    "barrier.  Per GPU and direction NVLink carries 2 x (G-1)/G of the live table bytes per sync: "
                    "inbound = the peers' replicas of the rows this rank owns (read responses) + the means the "
                    "peers write into this replica; outbound is the mirror image (NVLink 5: 900 GB/s per direction)"}

    # ---- e2e: the named job in full through the reference-facing call with HOST buffers in and
    # out: CSR H2D + init + one epoch over every start node (sharded over the ranks, replicas
    # averaged every sync_interval chunks) + tables D2H (rank 0); max over ranks ----
    if not args.no_e2e:
        engine.close()
        del engine
        torch.cuda.synchronize(device)
        e2e_engine = Engine(cfg["model"], embedding_size=D, epochs=1, iterations=cfg["iterations"],
                            return_weight=cfg["return_weight"], explore_weight=cfg["explore_weight"],
                            chunk_walks=args.chunk_walks, device=local_rank, **COMMON, **OPTIONS)
        out0 = out1 = pinned = None
        if rank == 0:
            pinned = PinnedTables(n, D)
            out0, out1 = pinned.tables
        if world > 1:
            dist.barrier()
        begin = time.perf_counter()
        e2e_engine.load_csr(graph.indptr, graph.indices)
        e2e_load_s = time.perf_counter() - begin
        if world > 1:
            e2e_engine.fit_distributed(SEED, args.sync_interval, gather="rank0", table0=out0, table1=out1)
        else:
            e2e_engine.fit(SEED, out0, out1)
        e2e_engine.sync()
        e2e_s = time.perf_counter() - begin
        digest = e2e_engine.tables_digest()
        job_counters = e2e_engine.counters()
        stats = torch.tensor([e2e_s, float(job_counters["pairs"])], dtype=torch.float64, device=device)
        if world > 1:
            slowest = stats[:1].clone()
            dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
            dist.all_reduce(stats[1:], op=dist.ReduceOp.SUM)
            stats[0] = slowest[0]
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            assert all(d["bits"] == digests[0]["bits"] for d in digests), "replicas differ after the last exchange"
            assert all(d["non_finite"] == 0 for d in digests), "non-finite values in the tables"
        assert digest["non_finite"] == 0, "non-finite values in the tables"
        e2e_s, e2e_pairs = float(stats[0]), float(stats[1])
        chunks_per_gpu = -(-(cfg["iterations"] * n_src) // (chunk * world))
        h2d = graph.indptr.nbytes + graph.indices.nbytes + 12 * n  # + sources, alias table
        d2h = 2 * n * D * 4
        result["e2e"] = {"value": e2e_pairs / e2e_s, "unit": "pairs/s",
                         "h2d_bytes_per_step": h2d / chunks_per_gpu,
                         "d2h_bytes_per_step": d2h / chunks_per_gpu,
                         "seconds": e2e_s, "steps_per_gpu": chunks_per_gpu, "pairs": e2e_pairs,
                         "exchanges": getattr(e2e_engine, "exchange_count", 0),
                         "phases_s_rank0": {"load_csr": e2e_load_s, **getattr(e2e_engine, "timings", {})},
                         "tables_checked": "finite" + (", replicas bit-identical on all ranks" if world > 1 else ""),
                         "call": ("b2e_load_csr + b2e_fit" if world == 1 else
                                  "Engine.load_csr + Engine.fit_distributed on every rank, tables to the host on rank 0") +
                                 f" (host CSR in, host tables out; the whole named job: epochs=1, "
                                 f"iterations={cfg['iterations']}, {cfg['iterations'] * n_src} walks)"}
        e2e_engine.close()
        del out0, out1
        if pinned is not None:
            pinned.release()
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
    parser.add_argument("--config", default="auto", choices=["auto"] + sorted(CONFIGS))
    parser.add_argument("--chunk-walks", type=int, default=1 << 20)
    parser.add_argument("--sync-interval", type=int, default=4)
    parser.add_argument("--reference-walks", type=int, default=0)
    parser.add_argument("--reference-step-seconds", type=float, default=2.0)
    parser.add_argument("--cpu-budget", type=float, default=15.0)
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--shared-negatives", action="store_true",
                        help="SkipGram: one set of negatives per centre (opt-in mode, not the named workload)")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)
    name, note = choose_config(args.config)
    cfg = CONFIGS[name]
    if args.shared_negatives:
        if cfg["model"] != "SkipGram":
            parser.error("--shared-negatives is a SkipGram option")
        OPTIONS["shared_negatives"] = True
    if args.impl == "reference":
        run_reference(args, name, cfg, note)
    else:
        run_ours(args, name, cfg, note)


if __name__ == "__main__":
    main()
