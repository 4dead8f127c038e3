"""Small run for compute-sanitizer (memcheck / racecheck): every model, both step paths, forward, sensors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dm_control_b200 import testing_models as tm
from dm_control_b200.physics import BatchedPhysics
for name, nsub in (('cartpole', 1), ('cheetah', 2), ('humanoid', 3), ('quadruped', 2), ('pendulum_free', 2)):
  model = tm.load(name); B = 6
  q0, v0 = tm.initial_states(model, name, B, 0)
  phys = BatchedPhysics(model, batch=B)
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  g = torch.Generator(device='cuda').manual_seed(0)
  for _ in range(6):
    phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
  torch.cuda.synchronize()
  print(name, 'ok', bool(torch.isfinite(phys.data.qpos).all()), phys.data.warning.sum(0).tolist())
