"""The batched torch twins of reference utilities against the REFERENCE'S OWN implementation, imported unmodified from
/root/reference (tests/refshim) — not against a restatement. Skipped where the reference checkout is absent (GPU box);
the restatement-based tests (tests/test_rewards_and_env_cpu.py, tests/test_randomizers_cpu.py) run everywhere.

  * dm_control_b200/rewards.py            vs dm_control/utils/rewards.py:25-135            (every sigmoid, bounds, margins)
  * dm_control_b200/control.compute_n_steps vs dm_control/rl/control.py:168-194            (values and error cases)
(The reference's randomizers and task files themselves run unmodified on the engine in tests/test_reference_tasks.py.)
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refshim   # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason='/root/reference is not on this machine')

SIGMOIDS = ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic', 'tanh_squared')


@pytest.fixture(scope='module')
def ref():
  refshim.install()
  import dm_control.utils.rewards as ref_rewards
  import dm_control.rl.control as ref_control
  assert os.path.realpath(ref_rewards.__file__).startswith(os.path.realpath(refshim.REFERENCE))
  return ref_rewards, ref_control


@pytest.mark.parametrize('sigmoid', SIGMOIDS)
def test_tolerance_equals_the_reference(ref, sigmoid):
  from dm_control_b200 import rewards
  ref_rewards, _ = ref
  rs = np.random.RandomState(0)
  x = np.concatenate([rs.uniform(-6, 6, 400), [-1.0, 0.0, 0.5, 1.0, 2.0, 3.0]])
  for bounds, margin, vam in (((0.0, 0.0), 1.0, 0.1), ((-1.0, 2.0), 0.5, 0.3), ((1.4, float('inf')), 0.35, 0.1), ((3.0, 3.0), 3.0, 0.0 if sigmoid in ('cosine', 'linear', 'quadratic') else 0.05),
                              ((0.0, 1.0), 0.0, 0.1)):
    got = rewards.tolerance(torch.as_tensor(x), bounds=bounds, margin=margin, sigmoid=sigmoid, value_at_margin=vam).numpy()
    want = ref_rewards.tolerance(x, bounds=bounds, margin=margin, sigmoid=sigmoid, value_at_margin=vam)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)


def test_tolerance_errors_equal_the_reference(ref):
  from dm_control_b200 import rewards
  ref_rewards, _ = ref
  for kw in (dict(bounds=(1.0, 0.0)), dict(margin=-1.0), dict(margin=1.0, sigmoid='nope'), dict(margin=1.0, sigmoid='gaussian', value_at_margin=0.0),
             dict(margin=1.0, sigmoid='linear', value_at_margin=1.0)):
    with pytest.raises(ValueError) as a:
      ref_rewards.tolerance(np.array([0.5]), **kw)
    with pytest.raises(ValueError) as b:
      rewards.tolerance(torch.tensor([0.5], dtype=torch.float64), **kw)
    assert str(a.value) == str(b.value)


def test_compute_n_steps_equals_the_reference(ref):
  from dm_control_b200 import control
  _, ref_control = ref
  for ct, pt in ((0.025, 0.005), (0.03, 0.005), (0.02, 0.0025), (0.01, 0.01), (0.04, 0.002)):
    assert control.compute_n_steps(ct, pt) == ref_control.compute_n_steps(ct, pt)
  for ct, pt in ((0.005, 0.01), (0.0251, 0.005)):
    with pytest.raises(ValueError) as a:
      ref_control.compute_n_steps(ct, pt)
    with pytest.raises(ValueError) as b:
      control.compute_n_steps(ct, pt)
    assert str(a.value) == str(b.value)


# ---- the batched task layer (dm_control_b200/suite/*) against the reference's own task files ---------------------------
_TASK_CHILD = r'''
import os, sys, importlib, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests'); sys.path.insert(0, %(root)r + '/tests/emu')
import gpu_shim; gpu_shim.install()
import refshim; refshim.install()
import numpy as np, torch
from dm_control_b200 import suite as bsuite
dom, task, B = %(dom)r, %(task)r, 4
benv = bsuite.load(dom, task, batch=B, seed=2, outputs='all')
benv.reset()
nu = benv.physics.model.nu
g = np.random.RandomState(0)
mod = importlib.import_module('dm_control.suite.' + dom)            # /root/reference/dm_control/suite/<dom>.py, unmodified
renv = getattr(mod, task)(random=0)
renv.reset()
rphys, rtask = renv.physics, renv.task
worst = {}
for t in range(8):
  a = g.uniform(-1, 1, (B, nu))
  ts = benv.step(torch.as_tensor(a, device=benv.physics.device))
  d = benv.physics.data
  for e in range(B):
    with rphys.reset_context():                                       # the batched environment's state, on the reference-facing view
      rphys.data.qpos[:] = d.qpos[e].cpu().numpy(); rphys.data.qvel[:] = d.qvel[e].cpu().numpy()
      if benv.physics.model.na: rphys.data.act[:] = d.act[e].cpu().numpy()
    rphys.set_control(a[e])
    robs, rrew = rtask.get_observation(rphys), rtask.get_reward(rphys)   # the reference's own observation / reward code
    assert set(robs) == set(ts.observation), (sorted(robs), sorted(ts.observation))
    for k, v in robs.items():
      if k in ('force_torque', 'imu'):      # acceleration-stage sensors: after a step they belong to the previous mj_forward, not to a fresh one
        continue
      got = ts.observation[k][e].cpu().numpy().reshape(-1)
      worst[k] = max(worst.get(k, 0.0), float(np.abs(got - np.asarray(v).reshape(-1)).max()))
    worst['reward'] = max(worst.get('reward', 0.0), abs(float(ts.reward[e]) - float(rrew)))
print('RESULT', json.dumps(worst))
'''


@pytest.mark.timeout(900)
@pytest.mark.parametrize('dom,task', [('cartpole', 'swingup'), ('cheetah', 'run'), ('humanoid', 'run'), ('humanoid', 'stand'), ('quadruped', 'walk')])
def test_batched_tasks_equal_the_reference_task_code(dom, task):
  """Observations and rewards of the batched tasks after random-action steps == what the reference's own
  `suite/<domain>.py` task computes on the same state (B = 1 reference-facing view; kernels from the CPU emulation build)."""
  import json, subprocess
  r = subprocess.run([sys.executable, '-c', _TASK_CHILD % dict(root=ROOT, dom=dom, task=task)], env=dict(os.environ, B200MJ_EMULATE_GPU='1'),
                     capture_output=True, text=True, timeout=800)
  assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
  worst = json.loads(r.stdout.split('RESULT', 1)[1])
  assert 'reward' in worst and len(worst) >= 3, worst
  assert max(worst.values()) < 1e-12, worst


_LOOP_CHILD = r'''
import os, sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests'); sys.path.insert(0, %(root)r + '/tests/emu')
import gpu_shim; gpu_shim.install()
import refshim; refshim.install()
import numpy as np, torch
from dm_control_b200 import suite as bsuite
import dm_control.suite.cartpole as ref_cartpole
TL = 0.055                                       # 5.5 control steps of 0.01 s: the episode ends on the 6th
renv = ref_cartpole.balance(time_limit=TL, random=0)
benv = bsuite.load('cartpole', 'balance', batch=3, seed=0, time_limit=TL)
rseq, bseq = [], []
ts = renv.reset(); rseq.append((int(ts.step_type), ts.discount))
tb = benv.reset(); bseq.append((int(tb.step_type[0]), None))
a = np.zeros(1)
for t in range(9):
  ts = renv.step(a); rseq.append((int(ts.step_type), None if ts.discount is None else float(ts.discount)))
  tb = benv.step(torch.zeros(3, 1, dtype=torch.float64, device=benv.physics.device))
  bseq.append((int(tb.step_type[1]), None if tb.discount is None else float(tb.discount[1])))
print('RESULT', json.dumps(dict(ref=rseq, batched=bseq)))
'''


@pytest.mark.timeout(900)
def test_environment_loop_equals_the_reference_loop():
  """step_type / discount sequence across a time limit and the automatic reset that follows: `BatchedEnvironment`
  (dm_control_b200/control.py) next to the reference's `control.Environment` (rl/control.py:77-127) on cartpole:balance."""
  import json, subprocess
  r = subprocess.run([sys.executable, '-c', _LOOP_CHILD % dict(root=ROOT)], env=dict(os.environ, B200MJ_EMULATE_GPU='1'),
                     capture_output=True, text=True, timeout=800)
  assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
  out = json.loads(r.stdout.split('RESULT', 1)[1])
  ref, got = out['ref'], out['batched']
  assert [s for s, _ in ref].count(2) >= 1 and [s for s, _ in ref].count(0) >= 2      # an end and a restart were seen
  # identical up to and including the LAST step (same step count to the time limit, discount 1.0 there) ...
  k = [s for s, _ in ref].index(2)
  assert ref[:k + 1] == got[:k + 1], out
  # ... then the one documented difference of a lock-stepped batch: the reference's next call only resets and returns
  # FIRST (rl/control.py:101-102), the batched environment resets that environment in place AND takes the step (the other
  # environments of the batch cannot wait), so its sequence is the reference's without that FIRST entry
  assert ref[k + 1][0] == 0
  assert ref[k + 2:] == got[k + 1:len(ref) - 1], out
