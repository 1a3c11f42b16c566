# usage: bash scripts/prof_walk.sh <tag> <config> [env assignments...]   (ncu --set full on one walk launch)
tag=$1; cfg=$2; shift; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk -s 2 -c 1 -o gpurun_out/prof_walk_$tag -f python bench.py --config $cfg --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_walk_$tag.log 2>&1
tail -2 gpurun_out/ncu_walk_$tag.log
