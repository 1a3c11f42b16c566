#!/bin/bash
# r02f: full GPU suite (resident graphs, one-pass Walklets, exchange unroll), sanitizers, load-time
# breakdown and the ncu traffic capture at the headline shape
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/r02f_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02f_pytest_gpu.txt
( time timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_new_kernels.py ) > gpurun_out/r02f_memcheck.txt 2>&1
tail -4 gpurun_out/r02f_memcheck.txt
( time timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_new_kernels.py ) > gpurun_out/r02f_racecheck.txt 2>&1
tail -4 gpurun_out/r02f_racecheck.txt
B2E_LOAD_TIMING=1 bash scripts/prof_train.sh r02f_c5 C5 B2E_LOAD_TIMING=1
grep "b2e load" gpurun_out/ncu_train_r02f_c5.log | head -12
python profiles/summarize.py gpurun_out/prof_train_r02f_c5.ncu-rep > gpurun_out/prof_train_r02f_c5.txt 2>&1
grep -E "duration|dram__bytes_(read|write).sum  |issue_active|warps_active|long_scoreboard|hit_rate" gpurun_out/prof_train_r02f_c5.txt
