"""Batched cheetah run (reference: dm_control/suite/cheetah.py)."""
from __future__ import annotations

import collections

import torch

from .. import control, rewards, testing_models
from ..physics import BatchedPhysics
from . import base

_DEFAULT_TIME_LIMIT = 10
_RUN_SPEED = 10
OUTPUTS = ('sensordata',)


class Physics(BatchedPhysics):

  def speed(self):
    """Horizontal speed (cheetah.py:55-57): sensordata['torso_subtreelinvel'][0]."""
    a = int(self.model.sensor_adr[self.model.names['sensor']['torso_subtreelinvel']])
    return self.data.sensordata[:, a]


class Cheetah(base.Task):

  def initialize_episode(self, physics, env_mask):
    """Random limited joints, then 200 settle steps, then time = 0 (cheetah.py:63-76)."""
    m = physics.model
    assert m.nq == m.njnt
    gen = self.generator(physics)
    physics.reset(env_mask=env_mask)
    lim = torch.as_tensor(m.jnt_limited == 1, device=physics.device)
    lo = torch.as_tensor(m.jnt_range[:, 0].copy(), device=physics.device)
    hi = torch.as_tensor(m.jnt_range[:, 1].copy(), device=physics.device)
    u = torch.rand(physics.batch, m.nq, generator=gen, device=physics.device, dtype=torch.float64)
    q = torch.where(lim, lo + u * (hi - lo), physics.data.qpos)
    saved = None
    if env_mask is not None:
      saved = (physics.get_state().clone(), physics.data.time.clone(), physics.data.qacc_warmstart.clone())
      physics.data.qpos[env_mask] = q[env_mask]
    else:
      physics.data.qpos.copy_(q)
    physics.step(nstep=200)
    if env_mask is not None:   # only the masked envs keep the settled state
      state, time, ws = physics.get_state().clone(), physics.data.time.clone(), physics.data.qacc_warmstart.clone()
      keep = ~env_mask
      state[keep], time[keep], ws[keep] = saved[0][keep], saved[1][keep], saved[2][keep]
      physics.set_state(state); physics.data.time.copy_(time); physics.data.qacc_warmstart.copy_(ws)
      physics.data.time[env_mask] = 0
      physics.forward()
    else:
      physics.data.time.zero_()

  def get_observation(self, physics):
    obs = collections.OrderedDict()
    obs['position'] = physics.data.qpos[:, 1:].clone()
    obs['velocity'] = physics.velocity().clone()
    return obs

  def get_reward(self, physics):
    return rewards.tolerance(physics.speed(), bounds=(_RUN_SPEED, float('inf')), margin=_RUN_SPEED, value_at_margin=0,
                             sigmoid='linear')


def _run(batch=1, seed=0, time_limit=_DEFAULT_TIME_LIMIT, **physics_kw):
  physics_kw.setdefault('outputs', OUTPUTS)
  physics = Physics(testing_models.load('cheetah'), batch=batch, **physics_kw)
  return control.BatchedEnvironment(physics, Cheetah(seed=seed), time_limit=time_limit)


TASKS = dict(run=_run)
