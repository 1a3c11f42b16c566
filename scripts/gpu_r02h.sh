#!/bin/bash
# r02h: GPU K3 + packed export: tests, then the headline shape with the load breakdown
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r02h_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02h_pytest_gpu.txt
( time B2E_LOAD_TIMING=1 timeout 1800 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02h_bench_c5.json 2> gpurun_out/r02h_bench_c5.err ) 2> gpurun_out/r02h_bench_c5.time
echo "rc=$?"; tail -3 gpurun_out/r02h_bench_c5.time; grep "b2e load" gpurun_out/r02h_bench_c5.err | tail -8
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench_c5.json"))
print("value %.4g" % d["value"], "e2e", json.dumps(d["e2e"]), "run", json.dumps(d["run"]))
PY
( time timeout 900 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_c4.json 2> gpurun_out/r02h_bench_c4.err ) 2> /dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench_c4.json"))
print("C4 value %.4g" % d["value"], "frac", d["roofline"]["frac"], "e2e", json.dumps(d["e2e"]))
PY
