#!/bin/bash
# r02k: the UBLKCP experiment (B2E_BULK=1) against the default SkipGram kernel, CBOW after the last diet
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_glove.py tests/test_gpu_sgns.py tests/test_gpu_fuzz.py -m gpu -q --maxfail=8 > gpurun_out/r02k_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02k_pytest_gpu.txt
timeout 900 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_c4.json 2> gpurun_out/r02k_bench_c4.err
for bulk in 0 1; do
  B2E_BULK=$bulk timeout 900 python bench.py --config C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_c2_bulk$bulk.json 2> gpurun_out/r02k_bench_c2_bulk$bulk.err
  B2E_BULK=$bulk timeout 900 python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_c3_bulk$bulk.json 2> gpurun_out/r02k_bench_c3_bulk$bulk.err
done
bash scripts/prof_train.sh r02k_c2_bulk1 C2 B2E_BULK=1
bash scripts/prof_train.sh r02k_c2_bulk0 C2 B2E_BULK=0
for t in prof_train_r02k_c2_bulk1 prof_train_r02k_c2_bulk0; do python profiles/summarize.py gpurun_out/$t.ncu-rep > gpurun_out/$t.txt 2>&1; done
python - <<'PY'
import json
for f in ("r02k_bench_c4", "r02k_bench_c2_bulk0", "r02k_bench_c2_bulk1", "r02k_bench_c3_bulk0", "r02k_bench_c3_bulk1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.5g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "kernel", d["roofline"]["kernel"], "ms", d["roofline"]["avg_launch_ms"])
    except Exception as ex:
        print(f, "failed", ex)
PY
grep -E "duration|inst_executed.sum|issue_active|l1tex__throughput|lts__throughput|registers|dram__bytes_(read|write).sum  " gpurun_out/prof_train_r02k_c2_bulk1.txt gpurun_out/prof_train_r02k_c2_bulk0.txt
