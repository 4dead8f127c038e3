"""Golden vectors from the reference's own PYTHON code (this container only: needs /root/reference), for machines that
do not have the checkout: tests/golden/reference_python_vectors.npz, checked by tests/test_reference_goldens.py.

  * `dm_control/utils/rewards.py: tolerance` on a fixed grid for every sigmoid and several (bounds, margin, value_at_margin);
  * the reference task files `dm_control/suite/{cartpole,cheetah,humanoid,quadruped}.py`: `get_observation` / `get_reward`
    evaluated on stored states — the states come from random-action rollouts of this engine (CPU emulation build of the
    kernels), the observation / reward arithmetic is the reference's, run unmodified on the B = 1 reference-facing view.

Run:  B200MJ_EMULATE_GPU=1 python tools/make_reference_goldens.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'emu')):
  sys.path.insert(0, p)
os.environ['B200MJ_EMULATE_GPU'] = '1'
import gpu_shim; gpu_shim.install()      # noqa: E402,E702
import refshim; refshim.install()        # noqa: E402,E702
import torch                             # noqa: E402
from dm_control_b200 import suite as bsuite      # noqa: E402

SIGMOIDS = ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic', 'tanh_squared')
REWARD_CASES = (((0.0, 0.0), 1.0, 0.1), ((-1.0, 2.0), 0.5, 0.3), ((1.4, float('inf')), 0.35, 0.1), ((0.0, 1.0), 0.0, 0.1))
TASKS = (('cartpole', 'swingup'), ('cartpole', 'balance'), ('cheetah', 'run'), ('humanoid', 'stand'), ('humanoid', 'run'), ('quadruped', 'walk'))
SKIP_KEYS = ('force_torque', 'imu')      # acceleration-stage sensors: not a function of (qpos, qvel, act, ctrl) alone after a step


def main():
  import dm_control.utils.rewards as ref_rewards
  out = {}
  x = np.concatenate([np.linspace(-6, 6, 241), [0.5, 1.4, 1.75, 3.0]])
  out['rewards_x'] = x
  for s in SIGMOIDS:
    for k, (bounds, margin, vam) in enumerate(REWARD_CASES):
      out[f'rewards_{s}_{k}'] = ref_rewards.tolerance(x, bounds=bounds, margin=margin, sigmoid=s, value_at_margin=vam)
  for dom, task in TASKS:
    B = 6
    benv = bsuite.load(dom, task, batch=B, seed=4, outputs='all')
    benv.reset()
    m = benv.physics.model
    g = np.random.RandomState(1)
    mod = importlib.import_module('dm_control.suite.' + dom)
    renv = getattr(mod, task)(random=0)
    renv.reset()
    rphys, rtask = renv.physics, renv.task
    for _ in range(12):      # a short random-action rollout of the batched environment: states off the reset manifold, in contact
      a = g.uniform(-1, 1, (B, m.nu))
      benv.step(torch.as_tensor(a, device=benv.physics.device))
    d = benv.physics.data
    qpos, qvel = d.qpos.cpu().numpy().copy(), d.qvel.cpu().numpy().copy()
    act = d.act.cpu().numpy().copy() if m.na else np.zeros((B, 0))
    obs_rows, rew = [], []
    keys = None
    for e in range(B):
      with rphys.reset_context():
        rphys.data.qpos[:] = qpos[e]; rphys.data.qvel[:] = qvel[e]
        if m.na: rphys.data.act[:] = act[e]
      rphys.set_control(a[e])
      robs = rtask.get_observation(rphys)
      keys = [k for k in robs if k not in SKIP_KEYS]
      obs_rows.append(np.concatenate([np.asarray(robs[k], dtype=np.float64).reshape(-1) for k in keys]))
      rew.append(float(rtask.get_reward(rphys)))
    tag = f'task_{dom}_{task}'
    out[tag + '_qpos'], out[tag + '_qvel'], out[tag + '_act'], out[tag + '_ctrl'] = qpos, qvel, act, a
    out[tag + '_obs'], out[tag + '_reward'] = np.stack(obs_rows), np.array(rew)
    out[tag + '_keys'] = np.array(keys)
    print(tag, 'obs', out[tag + '_obs'].shape, 'reward', np.round(rew, 4))
  path = os.path.join(ROOT, 'tests', 'golden', 'reference_python_vectors.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
