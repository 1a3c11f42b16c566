#!/bin/bash
# r02w: ncu capture of the shared-negative kernel on C3, its C5 rate, then the whole GPU suite + smoke on the final tree
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skipgram_shared_kernel -s 1 -c 1 \
  -o gpurun_out/prof_train_r02w_shared_c3 -f python bench.py --config C3 --shared-negatives --steps 2 --warmup 1 \
  --chunk-walks 131072 --no-e2e --no-cpu-baseline > gpurun_out/ncu_train_r02w.log 2>&1
tail -2 gpurun_out/ncu_train_r02w.log
timeout 600 python bench.py --shared-negatives --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02w_bench_C5_shared.json 2> gpurun_out/r02w_bench_C5_shared.err
echo "C5 shared rc=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/r02w_bench_C5_shared.json"))
    print(r["config"]["name"], "shared: value %.4g pairs/s, frac %.3f, sgd ms %.2f" % (r["value"], r["roofline"]["frac"], r["roofline"]["avg_launch_ms"]))
except Exception as error:
    print("no C5 result:", error)
PY
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02w_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02w_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
