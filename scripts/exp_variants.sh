timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log; grep -E "AUROC|oracle \[|held-out|oracle loss" gpurun_out/pytest_gpu.log
