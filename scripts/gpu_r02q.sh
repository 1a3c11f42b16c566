#!/bin/bash
# r02q: confirmation of the final tree: full GPU suite, smoke, the driver's two default commands
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r02q_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02q_reference_default.json 2> gpurun_out/r02q_reference_default.err ) 2> gpurun_out/r02q_reference_default.time
echo "reference rc=$?"; tail -3 gpurun_out/r02q_reference_default.time
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02q_bench_default.json 2> gpurun_out/r02q_bench_default.err ) 2> gpurun_out/r02q_bench_default.time
echo "ours rc=$?"; tail -3 gpurun_out/r02q_bench_default.time
python - <<'PY'
import json
for f in ("r02q_bench_default", "r02q_reference_default"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    e = d.get("e2e") or {}
    print(f, d["config"]["name"], "value %.5g" % d["value"], "frac", (d.get("roofline") or {}).get("frac"), "e2e %.5g" % e.get("value", 0), "cpu", (d.get("cpu_baseline") or {}).get("value"))
a, b = (json.load(open(f"gpurun_out/{f}.json"))["config"] for f in ("r02q_bench_default", "r02q_reference_default"))
print("same config:", a == b)
PY
