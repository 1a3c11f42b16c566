#!/bin/bash
# r02y: shared-negative kernel with the duplicate-free fast path (four CTAs per SM by default): rates,
# then the whole GPU suite + smoke on the final tree
mkdir -p gpurun_out
timeout 400 python bench.py --config C3 --shared-negatives --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02y_bench_C3_shared.json 2> gpurun_out/r02y_bench_C3_shared.err
echo "C3 rc=$?"
timeout 500 python bench.py --shared-negatives --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02y_bench_C5_shared.json 2> gpurun_out/r02y_bench_C5_shared.err
echo "C5 rc=$?"
python - <<'PY'
import json, glob
for path in sorted(glob.glob("gpurun_out/r02y_bench_*.json")):
    try:
        r = json.load(open(path))
        print(path, r["config"]["name"], "value %.4g pairs/s, frac %.3f, sgd ms %.2f traffic %s" % (r["value"], r["roofline"]["frac"], r["roofline"]["avg_launch_ms"], r["roofline"]["traffic"]))
    except Exception as error:
        print(path, "no result:", error)
PY
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02y_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02y_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
