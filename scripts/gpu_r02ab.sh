#!/bin/bash
# r02ab (2 GPUs): C5 with shared negatives, replicas averaged every 4 steps by the exchange kernel
mkdir -p gpurun_out
timeout 95 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --shared-negatives --no-e2e --no-cpu-baseline --steps 8 --warmup 3 \
  > gpurun_out/r02ab_bench_C5_shared_2gpu.json 2> gpurun_out/r02ab_bench_C5_shared_2gpu.err
echo "rc=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/r02ab_bench_C5_shared_2gpu.json"))
    print(r["config"]["name"], "n_gpus", r["n_gpus"], "value %.4g pairs/s" % r["value"], r.get("exchange"))
except Exception as error:
    print("no result:", error)
PY
tail -3 gpurun_out/r02ab_bench_C5_shared_2gpu.err
