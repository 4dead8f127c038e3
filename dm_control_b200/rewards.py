"""Batched torch twin of `dm_control/utils/rewards.py:25-135` (`tolerance` + sigmoids); same names and errors."""
from __future__ import annotations

import math

import torch

_DEFAULT_VALUE_AT_MARGIN = 0.1


def _sigmoids(x, value_at_1, sigmoid):
  if sigmoid in ('cosine', 'linear', 'quadratic'):
    if not 0 <= value_at_1 < 1:
      raise ValueError('`value_at_1` must be nonnegative and smaller than 1, got {}.'.format(value_at_1))
  else:
    if not 0 < value_at_1 < 1:
      raise ValueError('`value_at_1` must be strictly between 0 and 1, got {}.'.format(value_at_1))
  if sigmoid == 'gaussian':
    scale = math.sqrt(-2 * math.log(value_at_1))
    return torch.exp(-0.5 * (x * scale) ** 2)
  elif sigmoid == 'hyperbolic':
    scale = math.acosh(1 / value_at_1)
    return 1 / torch.cosh(x * scale)
  elif sigmoid == 'long_tail':
    scale = math.sqrt(1 / value_at_1 - 1)
    return 1 / ((x * scale) ** 2 + 1)
  elif sigmoid == 'reciprocal':
    scale = 1 / value_at_1 - 1
    return 1 / (abs(x) * scale + 1)
  elif sigmoid == 'cosine':
    scale = math.acos(2 * value_at_1 - 1) / math.pi
    scaled_x = x * scale
    return torch.where(abs(scaled_x) < 1, (1 + torch.cos(math.pi * scaled_x)) / 2, torch.zeros_like(x))
  elif sigmoid == 'linear':
    scaled_x = x * (1 - value_at_1)
    return torch.where(abs(scaled_x) < 1, 1 - scaled_x, torch.zeros_like(x))
  elif sigmoid == 'quadratic':
    scaled_x = x * math.sqrt(1 - value_at_1)
    return torch.where(abs(scaled_x) < 1, 1 - scaled_x ** 2, torch.zeros_like(x))
  elif sigmoid == 'tanh_squared':
    scale = math.atanh(math.sqrt(1 - value_at_1))
    return 1 - torch.tanh(x * scale) ** 2
  else:
    raise ValueError('Unknown sigmoid type {!r}.'.format(sigmoid))


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid='gaussian', value_at_margin=_DEFAULT_VALUE_AT_MARGIN):
  """1 inside `bounds`, sigmoidal fall-off outside (see the reference docstring, utils/rewards.py:93-135)."""
  lower, upper = bounds
  if lower > upper:
    raise ValueError('Lower bound must be <= upper bound.')
  if margin < 0:
    raise ValueError('`margin` must be non-negative.')
  x = torch.as_tensor(x)
  in_bounds = (lower <= x) & (x <= upper)
  one = torch.ones_like(x)
  if margin == 0:
    return torch.where(in_bounds, one, torch.zeros_like(x))
  d = torch.where(x < lower, lower - x, x - upper) / margin
  return torch.where(in_bounds, one, _sigmoids(d, value_at_margin, sigmoid))
