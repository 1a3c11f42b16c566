#!/bin/bash
# r02g (8 GPUs): the headline shape at N=8 as the driver's scaling run will launch it
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q > gpurun_out/r02g_pytest_8gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02g_pytest_8gpu.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 8 --warmup 2 > gpurun_out/r02g_bench_c5_8gpu.json 2> gpurun_out/r02g_bench_c5_8gpu.err ) 2> gpurun_out/r02g_bench_c5_8gpu.time
echo "rc=$?"; tail -3 gpurun_out/r02g_bench_c5_8gpu.time; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02g_bench_c5_8gpu.err | tail -5; head -c 7000 gpurun_out/r02g_bench_c5_8gpu.json; echo
free -g | head -2
