#!/bin/bash
# r02u (2 GPUs): the final tree on the multi-GPU path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/r02u_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02u_pytest_2gpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 8 --warmup 2 > gpurun_out/r02u_bench_c5_2gpu.json 2> gpurun_out/r02u_bench_c5_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02u_bench_c5_2gpu.json"))
print("value %.5g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "exchange ms", d["exchange"]["kernel_ms_per_sync"], "e2e %.5g" % d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["tables_checked"], "walk all gpus %.4g" % d["walk"]["steps_per_s_alone_all_gpus"])
PY
