timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log; grep "oracle loss" gpurun_out/pytest_gpu.log
fmt='
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        r=json.loads(line); print("pairs/s %.1fM  sgd_ms %.1f frac %.3f loss %.4f clocks %s walk %s" % (r["value"]/1e6, r["roofline"]["avg_launch_ms"], r["roofline"]["frac"], r["mean_pair_loss"], r["clocks"], r["walk"]))
    else: print(line.rstrip())
'
for V in 0 1; do for C in small_cbow small_n2v; do
echo "VARIANT=$V $C"
B2E_VARIANT=$V timeout 300 python bench.py --config $C --steps 5 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"
done; done
echo C2; timeout 300 python bench.py --steps 3 --warmup 1 --chunk-walks 262144 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"
