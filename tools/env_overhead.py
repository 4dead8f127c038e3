"""Where does a BatchedEnvironment.step go? kernel vs torch task layer (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dm_control_b200 import suite, testing_models as tm
B = 8192
env = suite.load('humanoid', 'run', batch=B, seed=0)
phys = env.physics; phys.check_errors = False
q0, v0 = tm.initial_states(phys.model, 'humanoid', B, 0)
phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward(); env._reset_next.zero_()
g = torch.Generator(device='cuda').manual_seed(0)
a = torch.empty(B, 21, dtype=torch.float64, device='cuda')
for _ in range(30):
  a.uniform_(-1, 1, generator=g); env._task.before_step(a, phys); phys.step(5)
def timeit(fn, n=20):
  torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
  for _ in range(n): fn()
  e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
print('physics.step(5)      %.3f ms' % timeit(lambda: phys.step(5)))
print('get_reward           %.3f ms' % timeit(lambda: env._task.get_reward(phys)))
print('get_observation      %.3f ms' % timeit(lambda: env._task.get_observation(phys)))
print('before_step          %.3f ms' % timeit(lambda: env._task.before_step(a, phys)))
print('uniform_             %.3f ms' % timeit(lambda: a.uniform_(-1, 1, generator=g)))
t0 = time.perf_counter(); 
for _ in range(20): env._task.get_reward(phys); env._task.get_observation(phys)
torch.cuda.synchronize(); print('host-side wall for reward+obs: %.3f ms' % ((time.perf_counter() - t0) / 20 * 1e3))
print('ws', phys.workspace_bytes(), 'epb', phys.envs_per_block(), 'warn', phys.data.warning.sum(0).tolist())
