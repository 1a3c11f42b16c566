timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log; grep "loss" gpurun_out/pytest_gpu.log | head
bash scripts/prof_train.sh pipe3 B2E_VARIANT=0
