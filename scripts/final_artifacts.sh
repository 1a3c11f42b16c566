# Regenerates the round's evidence on one B200 (about 10 GPU-minutes): gpu tests, smoke, both bench
# arms, the ncu launch list of the bench command, ncu --set full captures of the dominant kernels,
# the measurements of the SURVEY 8(f) rows and the memcheck run over the kernels added for them.
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; cat gpurun_out/final_bench_reference.json | cut -c1-300
python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err; cat gpurun_out/final_bench_c2.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1; tail -1 gpurun_out/final_ncu_bench.log | cut -c1-200
bash scripts/prof_train.sh final_c2 C2
timeout 300 python scripts/bench_next_rows.py > gpurun_out/next_rows.jsonl 2> gpurun_out/next_rows.err; tail -2 gpurun_out/next_rows.err
bash scripts/prof_glove.sh final
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_new_kernels.py > gpurun_out/sanitize.log 2>&1; tail -2 gpurun_out/sanitize.log
# then, here: python profiles/summarize.py gpurun_out/prof_train_final_c2.ncu-rep > profiles/rNN_skipgram_pipe_kernel_c2.txt  (etc.)
