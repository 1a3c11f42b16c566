#!/bin/bash
# r02s: sanitizers over the final kernels
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_new_kernels.py ) > gpurun_out/r02s_memcheck.txt 2>&1
tail -5 gpurun_out/r02s_memcheck.txt
( time timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_new_kernels.py ) > gpurun_out/r02s_racecheck.txt 2>&1
tail -5 gpurun_out/r02s_racecheck.txt
