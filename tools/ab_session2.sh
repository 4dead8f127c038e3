#!/bin/bash
# session A/B: base library (dm_control_b200/csrc/libb200mj_base.so, built from an earlier commit) against the tree's, CMU corridor + humanoid
for rep in 1 2; do
for v in base new; do
  so=$PWD/dm_control_b200/csrc/libb200mj.so; [ $v = base ] && so=$PWD/dm_control_b200/csrc/libb200mj_base.so
  B200MJ_SO=$so python tools/time_cmu_env.py $v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['label'], 'cmu ms', round(d['ms_per_step'],3), 'env/s', round(d['env_steps_per_s']), 'chk', d['qpos_checksum'], d['warnings'])"
  [ "$1" = cmu ] || B200MJ_SO=$so python bench.py --steps 30 --warmup 10 --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v humanoid value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3))"
done
done
