#!/bin/bash
# r02i (2 GPUs): exchange kernel with 4 rows per warp iteration; where the e2e seconds go at N > 1
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 2 > gpurun_out/r02i_bench_c5_2gpu.json 2> gpurun_out/r02i_bench_c5_2gpu.err ) 2> gpurun_out/r02i.time
echo "rc=$?"; tail -3 gpurun_out/r02i.time; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r02i_bench_c5_2gpu.err | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench_c5_2gpu.json"))
print("value %.4g" % d["value"], "ms/step", d["ms_per_step"], "sgd ms", d["roofline"]["avg_launch_ms"]); print("exchange", json.dumps(d["exchange"])); print("e2e", json.dumps(d["e2e"]))
PY
