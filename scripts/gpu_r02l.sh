#!/bin/bash
# r02l: SkipGram kernel at 5 CTAs per SM (B2E_SGD_OCC=5) on C2 / C3 / C5
mkdir -p gpurun_out
for cfg in C2 C3 C5; do
  B2E_SGD_OCC=5 timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02l_bench_${cfg}_occ5.json 2> gpurun_out/r02l_bench_${cfg}_occ5.err
done
timeout 900 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02l_bench_C5_occ4.json 2> gpurun_out/r02l_bench_C5_occ4.err
python - <<'PY'
import json
for f in ("r02l_bench_C2_occ5", "r02l_bench_C3_occ5", "r02l_bench_C5_occ5", "r02l_bench_C5_occ4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.5g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "ms", d["roofline"]["avg_launch_ms"], "loss", d["mean_pair_loss"])
    except Exception as ex:
        print(f, "failed", ex)
PY
