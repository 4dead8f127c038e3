"""Tiny driver for ncu: a few humanoid env-steps at batch B (default 2048)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dm_control_b200 import testing_models as tm
from dm_control_b200.physics import BatchedPhysics
name = sys.argv[1] if len(sys.argv) > 1 else 'humanoid'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
nsub = int(sys.argv[3]) if len(sys.argv) > 3 else 5
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
model = tm.load(name)
q0, v0 = tm.initial_states(model, name, B, 0)
phys = BatchedPhysics(model, batch=B, outputs=('xpos', 'xmat', 'subtree_com', 'sensordata'), full_final=False)
phys.check_errors = False
phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
g = torch.Generator(device='cuda').manual_seed(0)
# settle so the profiled steps see the steady-state contact load
for _ in range(40):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
torch.cuda.synchronize()
for _ in range(iters):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
torch.cuda.synchronize()
print('done', phys.data.warning.sum(0).tolist())
