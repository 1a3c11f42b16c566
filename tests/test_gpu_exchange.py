"""The exchange step of the data-parallel path (csrc/exchange.cu) on real hardware: replicas
averaged by one kernel per rank over peer memory.  The reference has no counterpart (SURVEY.md
8e); the contract is arithmetic: afterwards every replica holds, bit for bit, the float32 sum of
the replicas in rank order times 1 / world."""
import os
import socket

import numpy as np
import pytest

from embiggen_b200.engine import Engine

pytestmark = pytest.mark.gpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def expected_mean(tables):
    total = tables[0].copy()
    for t in tables[1:]:
        total = (total + t).astype(np.float32)
    return (total * np.float32(1.0 / len(tables))).astype(np.float32)


@pytest.mark.parametrize("world,D", [(2, 100), (3, 7), (4, 128), (8, 33)])
def test_local_replicas_are_averaged_exactly(er_graph, world, D):
    rng = np.random.default_rng(world)
    n = er_graph.get_number_of_nodes()
    replicas = [Engine("SkipGram", embedding_size=D) for _ in range(world)]
    try:
        inputs = []
        for engine in replicas:
            engine.load_csr(er_graph.indptr, er_graph.indices)
            t0 = rng.standard_normal((n, D)).astype(np.float32)
            t1 = rng.standard_normal((n, D)).astype(np.float32)
            engine.import_tables(t0, t1)
            inputs.append((t0, t1))
        for rank, engine in enumerate(replicas):
            engine.open_exchange_local(replicas, rank)
        for engine in replicas:
            engine.exchange_average()
        for engine in replicas:
            engine.sync()
        want0 = expected_mean([a for a, _ in inputs])
        want1 = expected_mean([b for _, b in inputs])
        digests = [engine.tables_digest() for engine in replicas]
        for engine in replicas:
            got0, got1 = engine.export_tables()
            assert np.array_equal(got0, want0) and np.array_equal(got1, want1)
        assert all(d["bits"] == digests[0]["bits"] and d["non_finite"] == 0 for d in digests)
        assert abs(digests[0]["sum"][0] - float(want0.astype(np.float64).sum())) < 1e-6 * n * D
        # a replica that drifts is seen by the digest
        t0, t1 = replicas[0].export_tables()
        t0[5, D - 1] = np.nextafter(t0[5, D - 1], np.float32(np.inf))
        replicas[0].import_tables(t0, t1)
        assert replicas[0].tables_digest()["bits"] != digests[1]["bits"]
        t1[0, 0] = np.nan
        replicas[0].import_tables(t0, t1)
        assert replicas[0].tables_digest()["non_finite"] == 1
    finally:
        for engine in replicas:
            engine.close()


def _ipc_worker(rank, world, port, out_dir, devices):
    """One process per replica, like one rank per GPU; the control plane is gloo so that two
    processes may share one GPU (NCCL refuses that), the data plane is CUDA IPC + the kernel."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from embiggen_b200.graph import erdos_renyi
    graph = erdos_renyi(3000, 20000, seed=5)
    kw = dict(embedding_size=24, walk_length=16, window_size=2, iterations=2, epochs=2,
              number_of_negative_samples=3, chunk_walks=512, device=devices[rank])
    with Engine("SkipGram", **kw) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        t0, t1, losses = engine.fit_distributed(9, sync_interval=2, gather="all")
        digest = engine.tables_digest()
    np.save(os.path.join(out_dir, f"t0_{rank}.npy"), t0)
    np.save(os.path.join(out_dir, f"t1_{rank}.npy"), t1)
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array(losses + [digest["non_finite"]]))
    with Engine("CBOW", **kw) as engine:  # rank-0 gather and the CBOW role order
        engine.load_csr(graph.indptr, graph.indices)
        c, x, _ = engine.fit_distributed(9, sync_interval=3, gather="rank0")
        assert (c is None) == (rank != 0)
        if rank == 0:
            raw0, raw1 = engine.export_tables()
            assert np.array_equal(c, raw1) and np.array_equal(x, raw0)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_fit_distributed_over_cuda_ipc(tmp_path, world):
    """Engine.fit_distributed end to end with one process per replica: IPC handles exchanged,
    peers mapped, walks sharded, replicas averaged by the kernel.  Uses one GPU per rank when the
    box has them, otherwise all ranks share GPU 0 (IPC across processes works either way)."""
    import torch
    import torch.multiprocessing as mp
    count = torch.cuda.device_count()
    devices = [r if count >= world else 0 for r in range(world)]
    mp.spawn(_ipc_worker, args=(world, free_port(), str(tmp_path), devices), nprocs=world, join=True)
    t0 = [np.load(tmp_path / f"t0_{r}.npy") for r in range(world)]
    t1 = [np.load(tmp_path / f"t1_{r}.npy") for r in range(world)]
    for r in range(1, world):
        assert np.array_equal(t0[0], t0[r]) and np.array_equal(t1[0], t1[r])
    meta = np.load(tmp_path / "meta_0.npy")
    assert meta[-1] == 0 and np.isfinite(t0[0]).all() and np.isfinite(t1[0]).all()
    # the averaged run lands where a single replica lands (same walks, same number of updates)
    from embiggen_b200.graph import erdos_renyi
    graph = erdos_renyi(3000, 20000, seed=5)
    with Engine("SkipGram", embedding_size=24, walk_length=16, window_size=2, iterations=2, epochs=2,
                number_of_negative_samples=3, chunk_walks=512) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        _, _, single = engine.fit(9)
    assert abs(meta[0] - single[0]) < 0.02 * single[0] and abs(meta[1] - single[1]) < 0.05 * single[1]
    assert not np.array_equal(t0[0], np.zeros_like(t0[0]))
