#!/bin/bash
# r02r: CBOW full-row variant A/B on C4 (bit-exact tests first), walk kernel ncu digest at C5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sgns.py tests/test_gpu_fuzz.py tests/test_golden.py -m gpu -q --maxfail=5 > gpurun_out/r02r_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02r_pytest_gpu.txt
timeout 900 python bench.py --config C4 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02r_bench_c4_full.json 2> gpurun_out/r02r_bench_c4_full.err
B2E_NO_FULL_ROWS=1 timeout 900 python bench.py --config C4 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02r_bench_c4_generic.json 2> gpurun_out/r02r_bench_c4_generic.err
bash scripts/prof_walk.sh r02r_c5 C5
python profiles/summarize.py gpurun_out/prof_walk_r02r_c5.ncu-rep > gpurun_out/r02r_walk_kernel_c5.txt 2>&1
python - <<'PY'
import json
for f in ("r02r_bench_c4_full", "r02r_bench_c4_generic"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value %.5g" % d["value"], "ms", d["roofline"]["avg_launch_ms"], "loss", d["mean_pair_loss"])
PY
grep -E "duration|dram__bytes_read.sum  |warps_active|registers" gpurun_out/r02r_walk_kernel_c5.txt
