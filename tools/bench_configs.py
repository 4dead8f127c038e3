"""Throughput of the other BASELINE.json configs (parity-test cases, not the bench line): env-steps/s on one GPU.

cartpole:swingup, cheetah:run B=4096, humanoid:run B=8192, quadruped:walk B=4096 run through BatchedEnvironment.step
(physics + reward + observation, random actions generated on device); the CMU humanoid (flat floor, nv=62) runs the
physics step only (its composer task layer is not built). Writes one JSON line per config.
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dm_control_b200 import suite, testing_models as tm
from dm_control_b200.physics import BatchedPhysics


def time_env(domain, task, B, warm=30, steps=40):
  env = suite.load(domain, task, batch=B, seed=0)
  env.reset()
  env.physics.check_errors = False
  nu = env.physics.model.nu
  g = torch.Generator(device='cuda').manual_seed(0)
  a = torch.empty(B, nu, dtype=torch.float64, device='cuda')
  env._graph_task_ops = True
  def one():
    a.uniform_(-1, 1, generator=g); env.task.before_step(a, env.physics); env.physics.step(env.n_sub_steps)
    env._reward_and_observation()
  for _ in range(warm): one()
  env.physics.data.warning.zero_()      # reset-time warnings (e.g. the quadruped's embedded start, quadruped.py:266-270) are not rollout warnings
  torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
  for _ in range(steps): one()
  e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e) / steps
  return dict(config=f'suite.{domain}:{task}', batch=B, n_sub_steps=env.n_sub_steps, ms_per_step=ms, env_steps_per_s=B / ms * 1e3,
              physics_steps_per_s=B * env.n_sub_steps / ms * 1e3, warnings=env.physics.data.warning.sum(0).tolist())


def time_cmu(B=2048, nsub=6, warm=15, steps=20):
  model = tm.load('cmu_humanoid')
  q0, v0 = tm.initial_states(model, 'cmu_humanoid', B, 0)
  phys = BatchedPhysics(model, batch=B, outputs=('xpos', 'xmat', 'sensordata', 'subtree_linvel'), full_final=False)
  phys.check_errors = False
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  g = torch.Generator(device='cuda').manual_seed(0)
  def one():
    phys.data.ctrl.uniform_(-1, 1, generator=g); phys.step(nsub)
  for _ in range(warm): one()
  torch.cuda.synchronize(); s = torch.cuda.Event(True); e = torch.cuda.Event(True); s.record()
  for _ in range(steps): one()
  e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e) / steps
  return dict(config='locomotion cmu_humanoid V2019 position-controlled, flat floor (physics step only)', batch=B, n_sub_steps=nsub,
              ms_per_step=ms, env_steps_per_s=B / ms * 1e3, physics_steps_per_s=B * nsub / ms * 1e3, warnings=phys.data.warning.sum(0).tolist())


if __name__ == '__main__':
  for args in (('cartpole', 'swingup', 4096), ('cheetah', 'run', 4096), ('humanoid', 'run', 8192), ('quadruped', 'walk', 4096)):
    print(json.dumps(time_env(*args)), flush=True)
  print(json.dumps(time_cmu()), flush=True)
