#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python -m pytest tests/test_quality.py tests/test_gpu_walks.py::test_load_rejects_malformed_csr_and_leaves_a_clean_handle -m gpu -q -s 2>&1 | grep -E "AUROC|passed|failed|oracle \[" >> gpurun_out/r02t_quality.txt
done
cat gpurun_out/r02t_quality.txt
