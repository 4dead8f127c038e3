"""Device-resident env-steps/s of the batched CMU corridor environment (config 5) — the same measurement as bench.py's
`configs` entry, alone, so that engine knobs (B200MJ_*) can be A/B'd in one GPU call: `python tools/time_cmu_env.py label`."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dm_control_b200 import locomotion

B = int(os.environ.get('CMU_BATCH', '2048'))
env = locomotion.load('cmu_humanoid_run_walls', batch=B, seed=3, device='cuda')
r = bench._time_env(env, int(os.environ.get('CMU_STEPS', '10')), 10, 5, env.physics.model.nu, 'cuda')
d = env.physics.data
nefc = d.nefc.reshape(-1)
r['nefc_le'] = {str(k): round(float((nefc <= k).double().mean()), 4) for k in (0, 6, 12, 32, 72)}
r.update(label=sys.argv[1] if len(sys.argv) > 1 else '', mean_nefc=float(d.nefc.double().mean()), max_nefc=int(d.nefc.max()),
         qpos_checksum=float(d.qpos.double().abs().sum()))
print(json.dumps(r), flush=True)
