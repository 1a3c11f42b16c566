"""The N > 1 path on CPU: world-size-2 (and 3) gloo process groups exercise the host logic of
the data-parallel driver -- walk-id sharding and replica averaging -- with the oracle standing
in for the kernels.  The reference has no multi-device path to mirror."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from embiggen_b200.engine import average_replicas, shard_chunks


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("per_epoch,capacity,world", [(10, 2, 2), (10640, 1 << 20, 8), (1001, 100, 3),
                                                       (7, 100, 8), (4096, 256, 4)])
def test_shards_partition_the_epoch(per_epoch, capacity, world):
    seen = []
    steps = None
    for rank in range(world):
        plan = list(shard_chunks(per_epoch, capacity, world, rank, base=5000))
        steps = len(plan) if steps is None else steps
        assert len(plan) == steps  # every rank makes the same number of steps (collectives line up)
        for first, count, stride in plan:
            assert stride == world and count <= capacity
            seen.extend(first + k * stride for k in range(count))
    assert sorted(seen) == list(range(5000, 5000 + per_epoch))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle
    from embiggen_b200.graph import erdos_renyi
    graph = erdos_renyi(300, 1500, seed=3)
    n, D, L, seed = graph.get_number_of_nodes(), 8, 12, 11
    per_epoch = int((np.diff(graph.indptr) > 0).sum())
    t0, t1 = oracle.init_tables(n, D, seed)  # identical replicas on every rank
    tables = [torch.from_numpy(t0), torch.from_numpy(t1)]
    walked = []
    for index, (first, count, stride) in enumerate(shard_chunks(per_epoch, 40, world, rank)):
        walks, _ = oracle.walks(graph.indptr, graph.indices, seed, first, count, L, 0.5, 2.0,
                                walk_id_stride=stride)
        walked.append((first, walks))
        if count:
            oracle.train("SkipGram", walks, t0, t1, seed, n, D, 2, 3, 0.05, first_walk=first,
                         walk_id_stride=stride)
        if (index + 1) % 2 == 0:
            average_replicas(tables)
    before = t0.copy()
    average_replicas(tables)
    gathered = [torch.zeros_like(tables[0]) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(before))
    np.save(os.path.join(out_dir, f"t0_{rank}.npy"), t0)
    np.save(os.path.join(out_dir, f"mean_{rank}.npy"), torch.stack(gathered).mean(0).numpy())
    np.save(os.path.join(out_dir, f"walks_{rank}.npy"), np.concatenate([w for _, w in walked]))
    np.save(os.path.join(out_dir, f"first_{rank}.npy"),
            np.concatenate([f + world * np.arange(len(w)) for f, w in walked]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_data_parallel_round_trip(tmp_path, world):
    mp.spawn(_worker, args=(world, free_port(), str(tmp_path)), nprocs=world, join=True)
    import oracle
    from embiggen_b200.graph import erdos_renyi
    graph = erdos_renyi(300, 1500, seed=3)
    per_epoch = int((np.diff(graph.indptr) > 0).sum())
    tables = [np.load(tmp_path / f"t0_{r}.npy") for r in range(world)]
    for r in range(1, world):  # every rank ends with the same replica ...
        assert np.array_equal(tables[0], tables[r])
    # ... which is the mean of the replicas before the last exchange
    assert np.allclose(tables[0], np.load(tmp_path / "mean_0.npy"), rtol=0, atol=1e-7)
    # the shards' walks are exactly the single-process walks of the same ids
    ids = np.concatenate([np.load(tmp_path / f"first_{r}.npy") for r in range(world)])
    walks = np.concatenate([np.load(tmp_path / f"walks_{r}.npy") for r in range(world)])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], np.arange(per_epoch))
    whole, _ = oracle.walks(graph.indptr, graph.indices, 11, 0, per_epoch, 12, 0.5, 2.0)
    assert np.array_equal(walks[order], whole)
    init0, _ = oracle.init_tables(300, 8, 11)
    assert not np.array_equal(tables[0], init0) and np.isfinite(tables[0]).all()
