"""Host-logic tests (no GPU): rewards twin vs the reference's numpy semantics; n_sub_steps arithmetic."""
import math

import numpy as np
import pytest
import torch

from dm_control_b200 import control, rewards


def _np_tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian', value_at_margin=0.1):
  """Independent numpy restatement of utils/rewards.py:93-135 for the sigmoids the suite uses."""
  lower, upper = bounds
  x = np.asarray(x, dtype=np.float64)
  inb = (lower <= x) & (x <= upper)
  if margin == 0:
    return np.where(inb, 1.0, 0.0)
  d = np.where(x < lower, lower - x, x - upper) / margin
  if sigmoid == 'gaussian':
    s = np.exp(-0.5 * (d * np.sqrt(-2 * np.log(value_at_margin))) ** 2)
  elif sigmoid == 'linear':
    sx = d * (1 - value_at_margin); s = np.where(abs(sx) < 1, 1 - sx, 0.0)
  elif sigmoid == 'quadratic':
    sx = d * np.sqrt(1 - value_at_margin); s = np.where(abs(sx) < 1, 1 - sx ** 2, 0.0)
  elif sigmoid == 'long_tail':
    s = 1 / ((d * np.sqrt(1 / value_at_margin - 1)) ** 2 + 1)
  return np.where(inb, 1.0, s)


@pytest.mark.parametrize('kw', [dict(bounds=(1.4, float('inf')), margin=0.35),
                                dict(bounds=(0.9, float('inf')), sigmoid='linear', margin=1.9, value_at_margin=0),
                                dict(margin=1, value_at_margin=0, sigmoid='quadratic'),
                                dict(bounds=(10, float('inf')), margin=10, value_at_margin=0, sigmoid='linear'),
                                dict(bounds=(-.25, .25)), dict(margin=2), dict(margin=5, sigmoid='long_tail')])
def test_tolerance_matches_reference_semantics(kw):
  x = np.linspace(-15, 15, 601)
  got = rewards.tolerance(torch.as_tensor(x), **kw).numpy()
  np.testing.assert_allclose(got, _np_tolerance(x, **kw), rtol=1e-14, atol=1e-15)
  assert got.min() >= 0 and got.max() <= 1


def test_tolerance_errors():
  with pytest.raises(ValueError, match='Lower bound must be <= upper bound.'):
    rewards.tolerance(torch.zeros(1), bounds=(1, 0))
  with pytest.raises(ValueError, match='`margin` must be non-negative.'):
    rewards.tolerance(torch.zeros(1), margin=-1)
  with pytest.raises(ValueError, match='Unknown sigmoid type'):
    rewards.tolerance(torch.ones(1), margin=1, sigmoid='nope')
  with pytest.raises(ValueError, match='strictly between 0 and 1'):
    rewards.tolerance(torch.ones(1), margin=1, value_at_margin=0)


def test_compute_n_steps():
  # dm_control/rl/control_test.py:122-127 and the two error paths of control.py:168-194
  assert control.compute_n_steps(0.03, 0.005) == 6
  assert control.compute_n_steps(0.025, 0.005) == 5
  assert control.compute_n_steps(0.01, 0.01) == 1
  with pytest.raises(ValueError, match='cannot be smaller'):
    control.compute_n_steps(0.001, 0.005)
  with pytest.raises(ValueError, match='integer multiple'):
    control.compute_n_steps(0.026, 0.005)


# ---- BatchedEnvironment loop with a mock physics / task (the reference tests its loop the same way, --------------
# ---- dm_control/rl/control_test.py:33-50) ---------------------------------------------------------------------------
class _MockPhysics:
  legacy_step = True

  def __init__(self, batch):
    self.batch, self.device = batch, torch.device('cpu')
    self.calls = []

  def timestep(self):
    return 0.01

  def step(self, n):
    self.calls.append(('step', n))


class _MockTask:
  def __init__(self, physics):
    self.resets = []
    self._p = physics

  def initialize_episode(self, physics, env_mask):
    self.resets.append(None if env_mask is None else env_mask.clone())

  def before_step(self, action, physics):
    physics.calls.append(('before_step', None))

  def after_step(self, physics):
    physics.calls.append(('after_step', None))

  def get_reward(self, physics):
    return torch.full((physics.batch,), 0.5, dtype=torch.float64)

  def get_observation(self, physics):
    return dict(x=torch.zeros(physics.batch, 3, dtype=torch.float64))


def test_environment_call_order_and_time_limit():
  """rl/control.py:99-127: before_step -> physics.step(n_sub_steps) -> after_step; the step that reaches the time
  limit is LAST, the next call re-initialises exactly those environments; no reset-flag readback in between."""
  phys = _MockPhysics(4)
  task = _MockTask(phys)
  env = control.BatchedEnvironment(phys, task, time_limit=0.06, control_timestep=0.02)   # 3 control steps, 2 sub-steps
  assert env.n_sub_steps == 2 and env.control_timestep() == pytest.approx(0.02)
  ts = env.reset()
  assert ts.step_type.tolist() == [control.FIRST] * 4 and ts.reward is None and len(task.resets) == 1
  types = []
  for _ in range(5):
    ts = env.step(torch.zeros(4, 1))
    types.append(int(ts.step_type[0]))
    assert ts.reward.tolist() == [0.5] * 4 and ts.discount.tolist() == [1.0] * 4
  assert types == [control.MID, control.MID, control.LAST, control.MID, control.MID]
  assert phys.calls[:3] == [('before_step', None), ('step', 2), ('after_step', None)]
  assert len(task.resets) == 2 and task.resets[1].tolist() == [True] * 4     # one masked re-initialisation, after LAST
  # partial reset requested from outside (e.g. diverged environments): only those environments are re-initialised
  env._reset_next = torch.tensor([False, True, False, False]); env._count_ub = float('inf')
  env.step(torch.zeros(4, 1))
  assert task.resets[-1].tolist() == [False, True, False, False]
  assert env._step_count.tolist() == [3, 1, 3, 3]


def test_environment_first_step_without_reset_initialises():
  """control.py:102: stepping before reset() re-initialises (here: every environment, through the mask)."""
  phys = _MockPhysics(2)
  task = _MockTask(phys)
  env = control.BatchedEnvironment(phys, task)
  env.step(torch.zeros(2, 1))
  assert len(task.resets) == 1 and task.resets[0].tolist() == [True, True]
  env.step(torch.zeros(2, 1))
  assert len(task.resets) == 1


def test_environment_task_termination():
  """control.py:113-121: a task termination ends the episode with the task's discount; the time limit wins with 1."""
  phys = _MockPhysics(3)
  task = _MockTask(phys)
  nan = float('nan')
  script = [torch.tensor([nan, nan, nan]), torch.tensor([nan, 0.0, nan]), torch.tensor([0.25, nan, nan])]
  task.get_termination = lambda physics: script.pop(0).to(torch.float64)
  env = control.BatchedEnvironment(phys, task, time_limit=0.03, control_timestep=0.01)     # 3 control steps
  env.reset()
  ts = env.step(torch.zeros(3, 1))
  assert ts.step_type.tolist() == [control.MID] * 3 and ts.discount.tolist() == [1.0] * 3
  ts = env.step(torch.zeros(3, 1))
  assert ts.step_type.tolist() == [control.MID, control.LAST, control.MID] and ts.discount.tolist() == [1.0, 0.0, 1.0]
  ts = env.step(torch.zeros(3, 1))                  # env 1 was re-initialised (count 1); envs 0 and 2 reach the time limit
  assert task.resets[-1].tolist() == [False, True, False]
  assert ts.step_type.tolist() == [control.LAST, control.MID, control.LAST]
  assert ts.discount.tolist() == [1.0, 1.0, 1.0]    # time limit: discount 1 even where the task also terminated (0.25)
