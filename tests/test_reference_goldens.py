"""tests/golden/reference_python_vectors.npz — vectors produced by the reference's own Python code in the build container
(tools/make_reference_goldens.py: `utils/rewards.py` and the task files `suite/{cartpole,cheetah,humanoid,quadruped}.py`,
imported unmodified) — against this repo's torch twins. Needs no reference checkout: runs anywhere, including the GPU box.

  * rewards: `dm_control_b200.rewards.tolerance` on the stored grid;
  * tasks: the batched tasks' observations / rewards on the stored states (kernels from the CPU emulation build, in a child
    process; `run_tasks_on_stored_states()` is the same body for a machine with a GPU).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden', 'reference_python_vectors.npz')
SIGMOIDS = ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic', 'tanh_squared')
REWARD_CASES = (((0.0, 0.0), 1.0, 0.1), ((-1.0, 2.0), 0.5, 0.3), ((1.4, float('inf')), 0.35, 0.1), ((0.0, 1.0), 0.0, 0.1))
TASKS = (('cartpole', 'swingup'), ('cartpole', 'balance'), ('cheetah', 'run'), ('humanoid', 'stand'), ('humanoid', 'run'), ('quadruped', 'walk'))


@pytest.mark.parametrize('sigmoid', SIGMOIDS)
def test_reward_twin_on_reference_vectors(sigmoid):
  from dm_control_b200 import rewards
  z = np.load(GOLD)
  x = torch.as_tensor(z['rewards_x'])
  for k, (bounds, margin, vam) in enumerate(REWARD_CASES):
    got = rewards.tolerance(x, bounds=bounds, margin=margin, sigmoid=sigmoid, value_at_margin=vam).numpy()
    np.testing.assert_allclose(got, z[f'rewards_{sigmoid}_{k}'], rtol=1e-13, atol=1e-15)


def run_tasks_on_stored_states(device=None):
  """-> worst |batched twin - reference vector| per task (called in-process on a GPU, in an emulation child on CPU)."""
  from dm_control_b200 import suite as bsuite
  z = np.load(GOLD)
  out = {}
  for dom, task in TASKS:
    tag = f'task_{dom}_{task}'
    qpos, qvel, act, ctrl = z[tag + '_qpos'], z[tag + '_qvel'], z[tag + '_act'], z[tag + '_ctrl']
    B = qpos.shape[0]
    env = bsuite.load(dom, task, batch=B, seed=0, outputs='all')
    env.reset()
    phys = env.physics
    d = phys.data
    d.qpos.copy_(torch.as_tensor(qpos)); d.qvel.copy_(torch.as_tensor(qvel))
    if act.shape[1]:
      d.act.copy_(torch.as_tensor(act))
    phys.set_control(torch.as_tensor(ctrl, device=phys.device))
    phys.forward()
    obs, rew = env.task.get_observation(phys), env.task.get_reward(phys)
    keys = [str(k) for k in z[tag + '_keys']]
    got = torch.cat([obs[k].reshape(B, -1) for k in keys], dim=1).cpu().numpy()
    out[tag] = max(float(np.abs(got - z[tag + '_obs']).max()), float(np.abs(rew.cpu().numpy() - z[tag + '_reward']).max()))
  return out


@pytest.mark.timeout(900)
def test_batched_tasks_on_reference_vectors_emulated():
  code = ("import os, sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r);"
          "import gpu_shim; gpu_shim.install();"
          "import test_reference_goldens as t; print('RESULT', json.dumps(t.run_tasks_on_stored_states()))") % (
              ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'emu'))
  r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, B200MJ_EMULATE_GPU='1'), capture_output=True, text=True, timeout=800)
  assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
  worst = json.loads(r.stdout.split('RESULT', 1)[1])
  assert len(worst) == len(TASKS) and max(worst.values()) < 1e-9, worst
