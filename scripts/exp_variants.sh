timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log; grep -E "AUROC|held-out" gpurun_out/pytest_gpu.log
