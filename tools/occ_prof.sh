for epb in 1 5; do
  B200MJ_ENVS_PER_BLOCK=$epb ncu --section WarpStateStats --section SchedulerStats --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:b200mj_step_kernel -s 42 -c 1 --csv --page raw python tools/prof_occ.py > gpurun_out/epb_$epb.csv 2>/dev/null
done
