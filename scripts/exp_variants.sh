fmt='
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        r=json.loads(line); print("pairs/s %.1fM  sgd_ms %.1f frac %.3f clocks %s" % (r["value"]/1e6, r["roofline"]["avg_launch_ms"], r["roofline"]["frac"], r["clocks"]["sm_mhz"]))
    else: print(line.rstrip())
'
for V in A B5 B4; do
cp scripts/ab/$V.so embiggen_b200/libb2e.so
echo "== $V"; timeout 300 python -m pytest tests/test_gpu_sgns.py -m gpu -q -x 2>&1 | tail -1
for C in C2 small_n2v; do echo "$V $C"; timeout 600 python bench.py --config $C --steps 4 --warmup 2 --chunk-walks 524288 --no-e2e --no-cpu-baseline 2>&1 | python -c "$fmt"; done
done
