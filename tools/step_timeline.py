"""Event timeline of one bench-style env step (diagnostic): where the non-kernel GPU time goes."""
import sys, os
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
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
packed = torch.empty(B, 69, dtype=torch.float64, device='cuda')
def pack(o, r):
  packed[:, :21] = o['joint_angles']; packed[:, 21] = o['head_height']; packed[:, 22:34] = o['extremities']
  packed[:, 34:37] = o['torso_vertical']; packed[:, 37:40] = o['com_velocity']; packed[:, 40:67] = o['velocity']; packed[:, 67] = r; packed[:, 68] = 1.0
names = ['flush', 'uniform', 'before_step', 'kernel', 'reward', 'obs', 'pack']
acc = {n: 0.0 for n in names}
for it in range(25):
  ev = [torch.cuda.Event(True) for _ in range(len(names) + 1)]
  ev[0].record(); flush.fill_(0.0)
  ev[1].record(); a.uniform_(-1, 1, generator=g)
  ev[2].record(); env._task.before_step(a, phys)
  ev[3].record(); phys.step(5)
  ev[4].record(); r = env._task.get_reward(phys)
  ev[5].record(); o = env._task.get_observation(phys)
  ev[6].record(); pack(o, r)
  ev[7].record(); torch.cuda.synchronize()
  if it >= 5:
    for k, n in enumerate(names): acc[n] += ev[k].elapsed_time(ev[k + 1]) / 20
print({k: round(v, 3) for k, v in acc.items()}, 'total', round(sum(acc.values()), 3))

# kernel time vs rollout age, and the cost of an nvidia-smi sampler running beside the loop
import subprocess, time
def run(n):
  s = torch.cuda.Event(True); e = torch.cuda.Event(True); ks = []
  torch.cuda.synchronize(); s.record()
  for _ in range(n):
    flush.fill_(0.0); a.uniform_(-1, 1, generator=g); env._task.before_step(a, phys)
    k0 = torch.cuda.Event(True); k1 = torch.cuda.Event(True); k0.record(); phys.step(5); k1.record(); ks.append((k0, k1))
    r = env._task.get_reward(phys); o = env._task.get_observation(phys); pack(o, r)
  e.record(); torch.cuda.synchronize()
  return s.elapsed_time(e) / n, sum(a_.elapsed_time(b_) for a_, b_ in ks) / n
for label in ('no sampler', 'nvidia-smi -lms 200', 'no sampler', 'nvidia-smi -lms 200'):
  p = None
  if 'smi' in label:
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm', '--format=csv,noheader', '-lms', '200'], stdout=subprocess.DEVNULL)
    time.sleep(0.5)
  tot, ker = run(20)
  if p: p.terminate()
  print(f'{label:22s} step {tot:.3f} ms  kernel {ker:.3f} ms  mean ncon {float(phys.data.ncon.float().mean()):.2f}')
