import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dm_control_b200 import testing_models as tm
from dm_control_b200.physics import BatchedPhysics
model = tm.load('humanoid'); B = 4096
q0, v0 = tm.initial_states(model, 'humanoid', B, 0)
phys = BatchedPhysics(model, batch=B, outputs=('xpos', 'xmat', 'subtree_com', 'sensordata'), full_final=False, nconmax=16, njmax=48)
phys.check_errors = False
phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
g = torch.Generator(device='cuda').manual_seed(0)
for _ in range(42):
  phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(5)
torch.cuda.synchronize()
