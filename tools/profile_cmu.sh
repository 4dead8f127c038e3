#!/bin/bash
# launch list of the CMU corridor environment's kernels (config 5): which kernel the 28 ms per control step go to
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:b200mj_ -s 150 -c 260 --csv --log-file gpurun_out/r2_cmu_launches.csv \
    python tools/run_cmu.py > gpurun_out/r2_cmu_under_ncu.log 2>&1
tail -2 gpurun_out/r2_cmu_under_ncu.log
