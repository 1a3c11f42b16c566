# usage: bash scripts/prof_glove.sh <tag>   (one GPU; ncu --set full on one production GloVe launch)
tag=$1
B2E_ROWS=glove timeout 600 ncu --set full --clock-control none --import-source on -k regex:glove_tile_kernel -s 1 -c 1 \
    -o gpurun_out/prof_glove_$tag -f python scripts/bench_next_rows.py > gpurun_out/ncu_glove_$tag.log 2>&1
tail -2 gpurun_out/ncu_glove_$tag.log
