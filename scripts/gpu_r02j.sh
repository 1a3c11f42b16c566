#!/bin/bash
# r02j: the closing single-GPU evidence: full suite, smoke, GloVe at the reference's defaults,
# complete bench lines for every north_star config, the driver's two default commands, the ncu
# launch list of the default command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r02j_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02j_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python - > gpurun_out/r02j_glove_defaults.txt 2>&1 <<'PY'
import time, numpy as np
from embiggen_b200.graph_gpu import rmat_gpu
from embiggen_b200.embedders import Node2VecGloVeB200
from embiggen_b200.engine import Engine
graph = rmat_gpu(20, 16_000_000, n=1_000_000, seed=3)
n_src = int((np.diff(graph.indptr) > 0).sum())
# the reference's defaults (node2vec_glove.py:8-30): embedding_size=100, walk_length=512, window_size=5, alpha=0.75
with Engine("GloVe", embedding_size=100, epochs=2, walk_length=512, window_size=5, iterations=1, return_weight=0.25,
            explore_weight=4.0, learning_rate=0.05, learning_rate_decay=0.9) as engine:
    engine.load_csr(graph.indptr, graph.indices)
    begin = time.perf_counter()
    c, x, losses = engine.fit(42)
    seconds = time.perf_counter() - begin
    counters = engine.counters()
slots = 2 * n_src * 512 * 5
print(f"GloVe, reference defaults (L=512, w=5, D=100), R-MAT 1M nodes / 16M edges, {n_src} start nodes: "
      f"{slots / 1e9:.2f} G key slots per epoch (limit of one sort: 2^31 = 2.15 G), 2 epochs in {seconds:.1f} s, "
      f"{counters['pairs'] / 1e9:.2f} G distinct triples trained in the last epoch, losses {losses}, finite "
      f"{bool(np.isfinite(c).all() and np.isfinite(x).all())}")
PY
cat gpurun_out/r02j_glove_defaults.txt | tail -3
for cfg in C2 C3 C4; do
  timeout 900 python bench.py --config $cfg --steps 10 --warmup 3 > gpurun_out/r02j_bench_$cfg.json 2> gpurun_out/r02j_bench_$cfg.err
  echo "$cfg rc=$?"
done
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_reference_default.json 2> gpurun_out/r02j_reference_default.err ) 2> gpurun_out/r02j_reference_default.time
echo "reference rc=$?"; tail -3 gpurun_out/r02j_reference_default.time
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02j_bench_default.json 2> gpurun_out/r02j_bench_default.err ) 2> gpurun_out/r02j_bench_default.time
echo "ours rc=$?"; tail -3 gpurun_out/r02j_bench_default.time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_c5.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/r02j_launches.err
python - <<'PY'
import json
for f in ("r02j_bench_C2", "r02j_bench_C3", "r02j_bench_C4", "r02j_bench_default", "r02j_reference_default"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        e = d.get("e2e") or {}
        print(f, d["config"]["name"], "value %.4g" % d["value"], "frac", (d.get("roofline") or {}).get("frac"), "e2e %.4g" % e.get("value", 0), "cpu", (d.get("cpu_baseline") or {}).get("value"), "walk", (d.get("walk") or {}).get("steps_per_s_alone"))
    except Exception as ex:
        print(f, "failed", ex)
PY
