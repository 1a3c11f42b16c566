timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_walks.py -m gpu -q -x > gpurun_out/pytest_fuzz.log 2>&1; tail -12 gpurun_out/pytest_fuzz.log
./scripts/microbench_rows 200000000 2>&1 | grep -E "stride 128 row 100|table" | tee gpurun_out/microbench_rows_100GB.log
