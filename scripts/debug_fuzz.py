import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle
from embiggen_b200.engine import Engine
from embiggen_b200.graph import erdos_renyi
case = {'graph': 'er', 'model': 'CBOW', 'D': 101, 'K': 5, 'w': 8, 'L': 16, 'rw': 2.0, 'ew': 4.0, 'lr': 0.025, 'alias': False, 'scale': True, 'normalize': True, 'seed': 1621167958129882765}
graph = erdos_renyi(2000, 12000, seed=7)
n = graph.get_number_of_nodes()
seed, D, L = case["seed"], case["D"], case["L"]
for w in (8, 7, 4):
  for n_walks in (1, 2, 10, 160):
    walks, _ = oracle.walks(graph.indptr, graph.indices, seed, 11, n_walks, L, case["rw"], case["ew"])
    t0, t1 = oracle.init_tables(n, D, seed)
    stats = oracle.train(case["model"], walks, t0, t1, seed, n, D, w, case["K"], case["lr"], 6.0, first_walk=11, indptr=graph.indptr, normalize_learning_rate_by_degree=True, scale_by_sqrt_dim=True)
    with Engine(case["model"], embedding_size=D, walk_length=L, window_size=w, iterations=1, number_of_negative_samples=case["K"], return_weight=case["rw"], explore_weight=case["ew"], use_scale_free_distribution=False, normalize_learning_rate_by_degree=True, scale_by_sqrt_dim=True, deterministic=True, chunk_walks=n_walks) as engine:
        engine.load_csr(graph.indptr, graph.indices)
        engine.init_tables(seed)
        engine.reset_counters()
        engine.walk_chunk(seed, 11, n_walks, 1, 0)
        engine.train_chunk(seed, 0, case["lr"])
        g0, g1 = engine.export_tables()
        c = engine.counters()
    print("variant", os.environ.get("B2E_VARIANT"), "w", w, "n_walks", n_walks, "oracle", stats["pairs"], stats["targets"], "gpu", c["pairs"], c["targets"], "tables equal", np.array_equal(g0, t0[:, :D]), np.array_equal(g1, t1[:, :D]))
