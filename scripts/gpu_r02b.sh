#!/bin/bash
# r02b: GPU test suite after the walk / exchange / load changes, then C3 with the new walk kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.txt 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02b_pytest_gpu.txt
tail -15 gpurun_out/r02b_pytest_gpu.txt
timeout 900 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err
echo "bench rc=$?"; tail -3 gpurun_out/r02b_bench_c3.err
B2E_NO_FILTER=1 timeout 600 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_c3_nofilter.json 2> gpurun_out/r02b_bench_c3_nofilter.err
B2E_NO_FOLD=1 timeout 600 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_c3_nofold.json 2> gpurun_out/r02b_bench_c3_nofold.err
python - <<'PY'
import json
for f in ("r02b_bench_c3", "r02b_bench_c3_nofilter", "r02b_bench_c3_nofold"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", d["value"], "walk", json.dumps(d["walk"]), "run", json.dumps(d["run"]), "e2e", d.get("e2e"))
    except Exception as e:
        print(f, "failed", e)
PY
