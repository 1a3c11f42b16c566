#!/bin/bash
# r02aa: the whole named job (C5, host CSR in -> one epoch over every start node -> host tables out) with shared negatives
mkdir -p gpurun_out
( time timeout 420 python bench.py --shared-negatives --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r02aa_bench_C5_shared_e2e.json 2> gpurun_out/r02aa_bench_C5_shared_e2e.err ) 2> gpurun_out/r02aa_time.txt
echo "rc=$?"; tail -3 gpurun_out/r02aa_time.txt
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/r02aa_bench_C5_shared_e2e.json"))
    print(r["config"]["name"], "value %.4g e2e %.4g" % (r["value"], r["e2e"]["value"]), r["e2e"].get("phases_s_rank0"), r["e2e"].get("seconds"))
except Exception as error:
    print("no result:", error)
PY
