timeout 600 python -m pytest tests/test_gpu_graph_build.py -m gpu -q -x > gpurun_out/pytest_graph.log 2>&1; tail -3 gpurun_out/pytest_graph.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; free -g | sed -n 2p; sleep 5; done ) > gpurun_out/c5_mem.log 2>&1 &
MON=$!
timeout 1200 python bench.py --config C5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err; echo "rc=$?"; cat gpurun_out/bench_C5.json | cut -c1-1500; tail -5 gpurun_out/bench_C5.err
kill $MON
sort -n gpurun_out/c5_mem.log | tail -2
