fmt='
import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        r=json.loads(line); print("pairs/s %.1fM  sgd_ms %.1f frac %.3f loss %.4f clocks %s walk %s" % (r["value"]/1e6, r["roofline"]["avg_launch_ms"], r["roofline"]["frac"], r["mean_pair_loss"], r["clocks"], r["walk"]))
    else: print(line.rstrip())
'
for C in C3 C4; do
echo "$C"; timeout 900 python bench.py --config $C --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/bench_$C.err | tee gpurun_out/bench_$C.json | python -c "$fmt"; tail -3 gpurun_out/bench_$C.err | head -2
done
