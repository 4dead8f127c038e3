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
