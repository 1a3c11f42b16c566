"""Quality (held-out edge AUROC) and throughput of the Hogwild launch as a function of the number
of concurrently trained walks, on a 20 000-node planted-partition graph (tuning of the automatic
`max_concurrent_walks`)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_quality import auroc
from embiggen_b200.engine import Engine
from embiggen_b200.graph import csr_from_edges

def planted(n, block, degree_in, degree_out, seed):
    rng = np.random.default_rng(seed)
    blocks = n // block
    src, dst = [], []
    for b in range(blocks):
        m = block * degree_in // 2
        a = rng.integers(0, block, m) + b * block
        c = rng.integers(0, block, m) + b * block
        src.append(a); dst.append(c)
    m = n * degree_out // 2
    src.append(rng.integers(0, n, m)); dst.append(rng.integers(0, n, m))
    src, dst = np.concatenate(src), np.concatenate(dst)
    keep = src != dst
    lo, hi = np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])
    keys = np.unique(lo * n + hi)
    return keys // n, keys % n

def holdout(src, dst, n, seed):
    rng = np.random.default_rng(seed)
    order = rng.permutation(len(src)); cut = int(0.8 * len(src))
    tr, te = order[:cut], order[cut:]
    existing = set((src * n + dst).tolist())
    def negatives(count):
        a = rng.integers(0, n, 2 * count); b = rng.integers(0, n, 2 * count)
        ok = (a != b) & ~np.isin(np.minimum(a, b) * n + np.maximum(a, b), src * n + dst)
        return np.stack([a[ok][:count], b[ok][:count]], 1)
    return (src[tr], dst[tr]), (src[te], dst[te]), negatives(len(tr)), negatives(len(te))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
src, dst = planted(n, 100, 10, 2, 1)
train_pos, test_pos, train_neg, test_neg = holdout(src, dst, n, 1)
graph = csr_from_edges(train_pos[0], train_pos[1], n)
print("nodes", n, "train edges", len(train_pos[0]))
kw = dict(embedding_size=64, walk_length=64, window_size=4, iterations=5, epochs=3, number_of_negative_samples=5,
          learning_rate=0.05, learning_rate_decay=0.9)
for model in ("SkipGram", "CBOW"):
    for cap in (n // 64, n // 16, n // 4, 1 << 20):
        with Engine(model, max_concurrent_walks=cap, **kw) as engine:
            engine.load_csr(graph.indptr, graph.indices)
            begin = time.perf_counter()
            c, x, losses = engine.fit(7)
            elapsed = time.perf_counter() - begin
            pairs = engine.counters()["pairs"] * kw["epochs"]
        a = auroc(np.hstack([c, x]), train_pos, test_pos, train_neg, test_neg)
        print(f"{model:8s} cap {cap:8d}  AUROC {a:.4f}  losses {np.round(losses, 4)}  {pairs / elapsed / 1e6:8.1f} M pairs/s (fit incl. export)")
