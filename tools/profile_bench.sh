#!/bin/bash
# Round-N profile of the bench command itself (run under gpurun). Outputs land in gpurun_out/; summaries are copied
# into profiles/ by tools/summarise_profiles.py.
#   1. launch list (every kernel with its device time; cold-cache, serialised: compare SHARES, not absolutes)
#   2. one `--set full` capture of the step kernel inside the same command
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b200mj_acc_kernel -s 42 -c 1 -o gpurun_out/prof_bench \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
# 3. DRAM traffic of every kernel of this library over two whole steps
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:b200mj -s 176 -c 44 --csv \
    --log-file gpurun_out/step_dram.csv python bench.py --steps 4 --warmup 8 --no-cpu > gpurun_out/bench_under_ncu_dram.log 2>&1
ls -la gpurun_out/
