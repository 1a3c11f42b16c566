#!/bin/bash
mkdir -p gpurun_out
for cfg in C2 C5; do
  B2E_SGD_OCC=3 timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02o_bench_${cfg}_occ3.json 2> gpurun_out/r02o_bench_${cfg}_occ3.err
done
python - <<'PY'
import json
for f in ("r02o_bench_C2_occ3", "r02o_bench_C5_occ3"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.5g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "ms", d["roofline"]["avg_launch_ms"])
    except Exception as ex:
        print(f, "failed", ex)
PY
