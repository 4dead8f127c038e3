"""Batched cartpole balance/swingup (reference: dm_control/suite/cartpole.py:130-225)."""
from __future__ import annotations

import collections
import math

import torch

from .. import control, rewards, testing_models
from ..physics import BatchedPhysics
from . import base

_DEFAULT_TIME_LIMIT = 10
OUTPUTS = ('xmat',)


class Physics(BatchedPhysics):

  def cart_position(self):
    return self.data.qpos[:, self.model.jnt_qposadr[self.model.names['joint']['slider']]]

  def angular_vel(self):
    return self.data.qvel[:, 1:]

  def pole_angle_cosine(self):
    return self.data.xmat[:, 2:, 8]                       # xmat[2:, 'zz']

  def bounded_position(self):
    poles = torch.stack([self.data.xmat[:, 2:, 8], self.data.xmat[:, 2:, 2]], dim=2).reshape(self.batch, -1)   # xmat[2:, ['zz', 'xz']].ravel()
    return torch.cat([self.cart_position()[:, None], poles], dim=1)


class Balance(base.Task):
  _CART_RANGE = (-.25, .25)
  _ANGLE_COSINE_RANGE = (.995, 1)

  def __init__(self, swing_up, sparse, seed=0):
    self._sparse, self._swing_up = sparse, swing_up
    super().__init__(seed)

  def initialize_episode(self, physics, env_mask):
    gen = self.generator(physics)
    B, nv, dev = physics.batch, physics.model.nv, physics.device
    physics.reset(env_mask=env_mask)
    def randn(*s): return torch.randn(*s, generator=gen, device=dev, dtype=torch.float64)
    def rand(*s): return torch.rand(*s, generator=gen, device=dev, dtype=torch.float64)
    q = torch.zeros(B, nv, dtype=torch.float64, device=dev)
    if self._swing_up:
      q[:, 0] = .01 * randn(B)
      q[:, 1] = math.pi + .01 * randn(B)
      if nv > 2:
        q[:, 2:] = .1 * randn(B, nv - 2)
    else:
      q[:, 0] = rand(B) * .2 - .1
      q[:, 1:] = rand(B, nv - 1) * .068 - .034
    v = 0.01 * randn(B, nv)
    if env_mask is None:
      physics.data.qpos.copy_(q); physics.data.qvel.copy_(v)
    else:
      physics.data.qpos[env_mask] = q[env_mask]; physics.data.qvel[env_mask] = v[env_mask]
    physics.after_reset(env_mask)

  def get_observation(self, physics):
    obs = collections.OrderedDict()
    obs['position'] = physics.bounded_position()
    obs['velocity'] = physics.velocity().clone()
    return obs

  def get_reward(self, physics):
    if self._sparse:
      cart_in_bounds = rewards.tolerance(physics.cart_position(), self._CART_RANGE)
      angle_in_bounds = rewards.tolerance(physics.pole_angle_cosine(), self._ANGLE_COSINE_RANGE).prod(dim=1)
      return cart_in_bounds * angle_in_bounds
    upright = (physics.pole_angle_cosine() + 1) / 2
    centered = (1 + rewards.tolerance(physics.cart_position(), margin=2)) / 2
    small_control = rewards.tolerance(physics.control(), margin=1, value_at_margin=0, sigmoid='quadratic')[:, 0]
    small_control = (4 + small_control) / 5
    small_velocity = rewards.tolerance(physics.angular_vel(), margin=5).min(dim=1).values
    small_velocity = (1 + small_velocity) / 2
    return upright.mean(dim=1) * small_control * small_velocity * centered


def _make(swing_up, sparse):
  def make(batch=1, seed=0, time_limit=_DEFAULT_TIME_LIMIT, **physics_kw):
    physics_kw.setdefault('outputs', OUTPUTS)
    physics = Physics(testing_models.load('cartpole'), batch=batch, **physics_kw)
    return control.BatchedEnvironment(physics, Balance(swing_up=swing_up, sparse=sparse, seed=seed), time_limit=time_limit)
  return make


TASKS = dict(balance=_make(False, False), balance_sparse=_make(False, True), swingup=_make(True, False),
             swingup_sparse=_make(True, True))
