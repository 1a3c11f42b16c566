#!/bin/bash
# Round-2 probe of the GPU box: host shape, whether the reference's wheel exists there, and
# the C5 (headline) bench with e2e as the code stood at the start of the round.
out=gpurun_out/r02a_probe.txt
{
echo "== host"; nproc; free -g; df -h /dev/shm | tail -1; lscpu | grep -E "Model name|Socket|Thread|Core" 
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,memory.total --format=csv
echo "== import ensmallen"; python -c "import ensmallen" 2>&1 | tail -1
echo "== import embiggen"; python -c "import embiggen" 2>&1 | tail -1
echo "== pip download ensmallen"; timeout 60 python -m pip download --no-deps -d /tmp/wheels "ensmallen>=0.8.94" 2>&1 | tail -2
echo "== pip download (wheelhouse)"; timeout 60 python -m pip download --no-index --find-links /opt/wheelhouse --no-deps -d /tmp/wheels ensmallen 2>&1 | tail -2
ls /opt/wheelhouse 2>/dev/null | grep -i -E "ensmallen|grape|embiggen" || echo "no ensmallen/grape/embiggen wheel in /opt/wheelhouse"
} > $out 2>&1
avail=$(free -g | awk '/Mem:/{print $7}')
echo "available GB: $avail" >> $out
if [ "$avail" -gt 300 ]; then
  timeout 1200 python bench.py --config C5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_c5_start.json 2> gpurun_out/r02a_bench_c5_start.err
  echo "bench rc=$?" >> $out
else
  timeout 900 python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/r02a_bench_c3_start.json 2> gpurun_out/r02a_bench_c3_start.err
  echo "bench(C3) rc=$?" >> $out
fi
free -g >> $out
cat $out
