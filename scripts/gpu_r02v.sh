#!/bin/bash
# r02v: the opt-in shared-negative SkipGram kernel: parity, sanitizers, C3 / C2 rates
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_shared_negatives.py -q -s > gpurun_out/r02v_pytest.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02v_pytest.txt; grep -h "AUROC\|oracle loss" gpurun_out/r02v_pytest.txt
( time timeout 240 compute-sanitizer --tool memcheck python scripts/sanitize_shared.py ) > gpurun_out/r02v_memcheck.txt 2>&1
tail -4 gpurun_out/r02v_memcheck.txt | head -3
( time timeout 300 compute-sanitizer --tool racecheck python scripts/sanitize_shared.py ) > gpurun_out/r02v_racecheck.txt 2>&1
tail -4 gpurun_out/r02v_racecheck.txt | head -3
timeout 500 python bench.py --config C3 --shared-negatives --no-e2e --steps 5 --warmup 3 > gpurun_out/r02v_bench_C3_shared.json 2> gpurun_out/r02v_bench_C3_shared.err
echo "C3 shared rc=$?"
timeout 300 python bench.py --config C2 --shared-negatives --no-e2e --steps 5 --warmup 3 > gpurun_out/r02v_bench_C2_shared.json 2> gpurun_out/r02v_bench_C2_shared.err
echo "C2 shared rc=$?"
python - <<'PY'
import json
for name in ("C3", "C2"):
    try:
        r = json.load(open(f"gpurun_out/r02v_bench_{name}_shared.json"))
        print(name, "shared: value %.4g pairs/s, frac %.3f, sgd ms %.2f, cpu %.4g" % (
            r["value"], r["roofline"]["frac"], r["roofline"]["avg_launch_ms"], (r.get("cpu_baseline") or {}).get("value", 0)))
    except Exception as error:
        print(name, "no result:", error)
PY
