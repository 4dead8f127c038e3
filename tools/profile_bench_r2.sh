#!/bin/bash
# Round-2 profile of the bench command (run under gpurun). Outputs land in gpurun_out/.
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_under_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b200mj_acc_tn_kernel -s 60 -c 1 -o gpurun_out/${TAG}_prof_acc \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_under_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b200mj_pos_kernel -s 20 -c 1 -o gpurun_out/${TAG}_prof_pos \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_under_ncu_full_pos.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
