#!/bin/bash
# Round-2 profile of the bench command (run under gpurun). Outputs land in gpurun_out/.
TAG=${1:-r2}
# launch list of this library's kernels over ~4 control steps in steady state (the task's reset loop comes first: skipped)
B200_BENCH_SETTLE=3 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:b200mj_ -s 260 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-configs > gpurun_out/${TAG}_under_ncu_launches.log 2>&1
B200_BENCH_SETTLE=3 ncu --set full --clock-control none --import-source on -k regex:b200mj_acc_tn_kernel -s 80 -c 1 -o gpurun_out/${TAG}_prof_acc \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/${TAG}_under_ncu_full.log 2>&1
B200_BENCH_SETTLE=3 ncu --set full --clock-control none --import-source on -k regex:b200mj_pos_ -s 30 -c 1 -o gpurun_out/${TAG}_prof_pos \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/${TAG}_under_ncu_full_pos.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
