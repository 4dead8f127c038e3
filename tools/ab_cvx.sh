#!/bin/bash
# convex pairs by the whole warp (B200MJ_CVX_WARP=1, default) against one pair per lane, same library: CMU corridor + quadruped
for rep in 1 2; do
for v in 1 0; do
  B200MJ_CVX_WARP=$v python tools/time_cmu_env.py cvx_warp=$v 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['label'], 'cmu ms', round(d['ms_per_step'],3), 'env/s', round(d['env_steps_per_s']), 'chk', d['qpos_checksum'])"
  B200MJ_CVX_WARP=$v python - <<PY
import torch, sys
sys.path.insert(0, '.')
import bench
from dm_control_b200 import suite
e = suite.load('quadruped', 'walk', batch=4096, seed=3, device='cuda')
r = bench._time_env(e, 20, 10, 5, e.physics.model.nu, 'cuda'); print('cvx_warp=$v quadruped env/s', round(r['env_steps_per_s']), 'ms', round(r['ms_per_step'], 3), 'chk', float(e.physics.data.qpos.abs().sum()))
PY
done
done
