timeout 1200 python -m pytest tests/test_gpu_sgns.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
fmt='
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        r=json.loads(line); print("pairs/s %.1fM  sgd_ms %.1f frac %.3f clocks %s walk_ms %.2f steps/s %.2fG" % (r["value"]/1e6, r["roofline"]["avg_launch_ms"], r["roofline"]["frac"], r["clocks"], r["walk"]["avg_launch_ms"], r["walk"]["steps_per_s_alone"]/1e9))
    else: print(line.rstrip())
'
echo C2; timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"
for SM in 1 0; do echo "small_n2v WALK_SM=$SM"; B2E_WALK_SM=$SM timeout 600 python bench.py --config small_n2v --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"; done
echo C4; timeout 600 python bench.py --config C4 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"
