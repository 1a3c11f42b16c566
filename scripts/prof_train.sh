# usage: bash scripts/prof_train.sh <tag> <config> [env assignments...]   (one GPU; ncu --set full on one SGD launch)
tag=$1; cfg=$2; shift; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:train_kernel\|pipe_kernel -s 1 -c 1 -o gpurun_out/prof_train_$tag -f python bench.py --config $cfg --steps 2 --warmup 1 --chunk-walks 131072 --no-e2e --no-cpu-baseline > gpurun_out/ncu_train_$tag.log 2>&1
tail -2 gpurun_out/ncu_train_$tag.log
