#!/bin/bash
# r02x: shared-negative kernel with K + 1 rows per stage; three against four CTAs per SM
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_shared_negatives.py -q > gpurun_out/r02x_pytest.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02x_pytest.txt
for occ in 3 4; do
  B2E_SGD_OCC=$occ timeout 400 python bench.py --config C3 --shared-negatives --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02x_bench_C3_shared_occ$occ.json 2> gpurun_out/r02x_bench_C3_shared_occ$occ.err
  echo "C3 occ $occ rc=$?"
done
B2E_SGD_OCC=4 timeout 500 python bench.py --shared-negatives --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02x_bench_C5_shared_occ4.json 2> gpurun_out/r02x_bench_C5_shared_occ4.err
echo "C5 occ 4 rc=$?"
python - <<'PY'
import json, glob
for path in sorted(glob.glob("gpurun_out/r02x_bench_*.json")):
    try:
        r = json.load(open(path))
        print(path, r["config"]["name"], "value %.4g pairs/s, frac %.3f, sgd ms %.2f" % (r["value"], r["roofline"]["frac"], r["roofline"]["avg_launch_ms"]))
    except Exception as error:
        print(path, "no result:", error)
PY
