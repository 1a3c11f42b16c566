#!/bin/bash
# r02m (2 GPUs): fit_distributed with the peers mapped under the first chunks; exchange rows per iteration
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_sgns.py::test_sink_centre_with_degree_normalised_learning_rate_stays_finite -m gpu -q > gpurun_out/r02m_pytest_2gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest_2gpu.txt
run() { tag=$1; shift; env "$@" timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 8 --warmup 2 $EXTRA > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "$tag rc=$?"; }
EXTRA="" run r02m_bench_c5_2gpu B2E_EXCHANGE_ROWS=0
EXTRA="--no-e2e" run r02m_bench_c5_2gpu_rows8 B2E_EXCHANGE_ROWS=8
EXTRA="--no-e2e" run r02m_bench_c5_2gpu_rows1 B2E_EXCHANGE_ROWS=1
python - <<'PY'
import json
for f in ("r02m_bench_c5_2gpu", "r02m_bench_c5_2gpu_rows8", "r02m_bench_c5_2gpu_rows1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.5g" % d["value"], "exchange", json.dumps(d["exchange"])[:420]); print("   e2e", json.dumps(d.get("e2e"))[:700])
    except Exception as ex:
        print(f, "failed", ex)
PY
