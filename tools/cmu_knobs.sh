#!/bin/bash
# CMU corridor config: engine knob sweep in one GPU call (tools/time_cmu_env.py per variant)
run() { env "$@" python tools/time_cmu_env.py "$*" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['label'], '| ms', round(d['ms_per_step'],3), 'env/s', round(d['env_steps_per_s']), d['nefc_le'], d['warnings'][:3])"; }
run A=0
run B200MJ_EPB_POS=6
run B200MJ_EPB_POS=5
run B200MJ_EPB_POS=4
run B200MJ_ACC_WARPS_B=6,3,3,1
run B200MJ_ACC_WARPS_B=5,3,3,1
run B200MJ_BUCKETS=12,32
