"""Occupancy experiment: humanoid env-step time vs resident warps/SM (padding dynamic smem per CTA)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from dm_control_b200 import testing_models as tm
from dm_control_b200.physics import BatchedPhysics
model = tm.load('humanoid'); B = 8192
q0, v0 = tm.initial_states(model, 'humanoid', B, 0)
phys = BatchedPhysics(model, batch=B, outputs=('xpos','xmat','subtree_com','sensordata'), full_final=False, nconmax=int(os.environ.get('NCON','16')), njmax=int(os.environ.get('NJ','48')))
phys.check_errors = False
phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
g = torch.Generator(device='cuda').manual_seed(0)
for _ in range(30):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(5)
torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); s.record()
for _ in range(10):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(5)
e.record(); torch.cuda.synchronize()
print('caps', os.environ.get('NCON'), os.environ.get('NJ'), 'warn', phys.data.warning.sum(0).tolist()[1:3], 'sync', os.environ.get('B200MJ_SYNC_LEVEL'), 'envs/block', os.environ.get('B200MJ_ENVS_PER_BLOCK'), phys.envs_per_block(), 'ws', phys.workspace_bytes(), 'ms', s.elapsed_time(e)/10)
'''
for pad, sl, ncon, nj in ((8, 1, 16, 48), (8, 1, 12, 40), (8, 1, 10, 32), (8, 1, 8, 24), (8, 0, 8, 24), (8, 2, 8, 24)):
  env = dict(os.environ, B200MJ_ENVS_PER_BLOCK=str(pad), B200MJ_SYNC_LEVEL=str(sl), NCON=str(ncon), NJ=str(nj))
  print(subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
