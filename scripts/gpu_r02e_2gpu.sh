#!/bin/bash
# r02e (2 GPUs): the exchange kernel over real NVLink peers, then bench.py at N=2 (C3, then the headline shape)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02e_topo.txt 2>&1; free -g | head -2; nproc
timeout 900 python -m pytest tests/test_gpu_exchange.py "tests/test_quality.py::test_replica_averaging_keeps_the_quality" -m gpu -q -s > gpurun_out/r02e_pytest_2gpu.txt 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02e_pytest_2gpu.txt
run() { tag=$1; shift; ( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err ) 2> gpurun_out/$tag.time; echo "$tag rc=$?"; tail -3 gpurun_out/$tag.time; tail -4 gpurun_out/$tag.err; head -c 6000 gpurun_out/$tag.json; echo; }
run r02e_bench_c3_2gpu --config C3 --steps 8 --warmup 2
run r02e_bench_c5_2gpu --steps 8 --warmup 2
free -g | head -2
