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
phys = BatchedPhysics(model, batch=B, outputs=('xpos','xmat','subtree_com','sensordata'), full_final=False, nconmax=16, njmax=48)
phys.check_errors = False
phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
g = torch.Generator(device='cuda').manual_seed(0)
for _ in range(30):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(5)
torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True); s.record()
for _ in range(10):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(5)
e.record(); torch.cuda.synchronize()
print('sync', os.environ.get('B200MJ_SYNC_LEVEL'), 'envs/block', os.environ.get('B200MJ_ENVS_PER_BLOCK'), phys.envs_per_block(), 'ws', phys.workspace_bytes(), 'ms', s.elapsed_time(e)/10)
'''
for pad, sl in ((5, 0), (5, 1), (5, 2), (5, 3), (4, 1), (4, 2), (3, 1), (2, 1)):
  env = dict(os.environ, B200MJ_ENVS_PER_BLOCK=str(pad), B200MJ_SYNC_LEVEL=str(sl))
  print(subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
