#!/bin/bash
mkdir -p gpurun_out
B2E_SGD_OCC=3 timeout 900 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02p_bench_C3_occ3.json 2> gpurun_out/r02p_bench_C3_occ3.err
B2E_SGD_OCC=2 timeout 900 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02p_bench_C5_occ2.json 2> gpurun_out/r02p_bench_C5_occ2.err
B2E_SGD_OCC=3 timeout 900 python bench.py --config C5 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02p_bench_C5_occ3.json 2> gpurun_out/r02p_bench_C5_occ3.err
B2E_SGD_OCC=4 timeout 900 python bench.py --config C5 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02p_bench_C5_occ4.json 2> gpurun_out/r02p_bench_C5_occ4.err
python - <<'PY'
import json
for f in ("r02p_bench_C3_occ3", "r02p_bench_C5_occ2", "r02p_bench_C5_occ3", "r02p_bench_C5_occ4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.5g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "ms", d["roofline"]["avg_launch_ms"], "clk", d["clocks"]["sm_mhz"])
    except Exception as ex:
        print(f, "failed", ex)
PY
