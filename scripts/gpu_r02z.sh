#!/bin/bash
# r02z: held-out-edge AUROC of the shared-negative estimator; ncu capture of its final kernel on C3
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_quality.py -q -s -k shared_negatives > gpurun_out/r02z_quality_shared.txt 2>&1
echo "pytest rc=$?"; grep -h "AUROC\|mean over\|passed\|failed" gpurun_out/r02z_quality_shared.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:skipgram_shared_kernel -s 1 -c 1 \
  -o gpurun_out/prof_train_r02z_shared_c3 -f python bench.py --config C3 --shared-negatives --steps 2 --warmup 1 \
  --chunk-walks 131072 --no-e2e --no-cpu-baseline > gpurun_out/ncu_train_r02z.log 2>&1
tail -1 gpurun_out/ncu_train_r02z.log
