"""Batched quadruped walk/run (reference: dm_control/suite/quadruped.py:130-356, `Move` task, flat floor).

Deviation recorded in DESIGN.md §7: ellipsoid/cylinder geoms collide with planes only (the convex-convex pairs —
torso ellipsoid or eye cylinders against leg capsules/spheres — are not in the narrow-phase subset).
"""
from __future__ import annotations

import collections
import math

import torch

from .. import control, rewards, testing_models
from ..physics import BatchedPhysics
from . import base

_DEFAULT_TIME_LIMIT = 20
_CONTROL_TIMESTEP = .02
_RUN_SPEED = 5
_WALK_SPEED = 0.5
_TOES = ['toe_front_left', 'toe_back_left', 'toe_back_right', 'toe_front_right']
OUTPUTS = ('xpos', 'xmat', 'sensordata', 'ncon')

_SENS_GYRO, _SENS_ACCEL, _SENS_FORCE, _SENS_TORQUE = 3, 1, 4, 5


class Physics(BatchedPhysics):

  def _cache(self):
    if not hasattr(self, '_q'):
      m, dev = self.model, self.device
      n = m.names
      hinge = [j for j in range(m.njnt) if m.jnt_type[j] == 3]
      def sens_idx(types):
        idx = []
        for s in range(m.nsensor):
          if int(m.sensor_type[s]) in types:
            idx += list(range(int(m.sensor_adr[s]), int(m.sensor_adr[s]) + int(m.sensor_dim[s])))
        return torch.tensor(idx, dtype=torch.int64, device=dev)
      vel = int(m.sensor_adr[n['sensor']['velocimeter']])
      self._q = dict(torso=n['body']['torso'],
                     hinge_q=torch.tensor([int(m.jnt_qposadr[j]) for j in hinge], dtype=torch.int64, device=dev),
                     hinge_v=torch.tensor([int(m.jnt_dofadr[j]) for j in hinge], dtype=torch.int64, device=dev),
                     toes=torch.tensor([n['body'][t] for t in _TOES], dtype=torch.int64, device=dev),
                     vel=vel, imu=sens_idx((_SENS_GYRO, _SENS_ACCEL)), ft=sens_idx((_SENS_FORCE, _SENS_TORQUE)))
    return self._q

  def torso_upright(self):
    return self.data.xmat[:, self._cache()['torso'], 8]

  def torso_velocity(self):
    a = self._cache()['vel']
    return self.data.sensordata[:, a:a + 3].clone()

  def egocentric_state(self):
    c = self._cache()
    return torch.cat([self.data.qpos.index_select(1, c['hinge_q']), self.data.qvel.index_select(1, c['hinge_v']), self.data.act], dim=1)

  def toe_positions(self):
    c = self._cache()
    frame = self.data.xmat[:, c['torso']].reshape(-1, 3, 3)
    rel = self.data.xpos.index_select(1, c['toes']) - self.data.xpos[:, c['torso']][:, None, :]
    return torch.bmm(rel, frame)

  def force_torque(self):
    return torch.arcsinh(self.data.sensordata.index_select(1, self._cache()['ft']))

  def imu(self):
    return self.data.sensordata.index_select(1, self._cache()['imu'])


def _upright_reward(physics, deviation_angle=0):
  deviation = math.cos(math.radians(deviation_angle))
  return rewards.tolerance(physics.torso_upright(), bounds=(deviation, float('inf')), sigmoid='linear', margin=1 + deviation,
                           value_at_margin=0)


class Move(base.Task):

  def __init__(self, desired_speed, seed=0):
    self._desired_speed = desired_speed
    super().__init__(seed)

  def initialize_episode(self, physics, env_mask):
    """Random orientation, then raise the root 1 cm at a time until contact-free (quadruped.py:250-279,328-339)."""
    gen = self.generator(physics)
    B, dev = physics.batch, physics.device
    todo = torch.ones(B, dtype=torch.bool, device=dev) if env_mask is None else env_mask.clone()
    physics.reset(env_mask=env_mask)
    quat = torch.randn(B, 4, generator=gen, device=dev, dtype=torch.float64)
    quat = quat / quat.norm(dim=1, keepdim=True)
    z = torch.zeros(B, dtype=torch.float64, device=dev)
    for _ in range(400):
      q = physics.data.qpos
      q[todo, 0] = 0.0; q[todo, 1] = 0.0
      q[todo, 2] = z[todo]
      q[todo, 3:7] = quat[todo]
      with physics.suppress_physics_errors():      # a full contact buffer while embedded is expected (quadruped.py:266-270)
        physics.after_reset(todo)
      todo = todo & (physics.data.ncon > 0)
      z = z + 0.01
      if not bool(todo.any()):
        break
    else:
      # quadruped.py:271-276: 'Failed to find a non-contacting configuration.' after its attempt budget
      raise RuntimeError('Failed to find a non-contacting configuration.')

  def get_observation(self, physics):
    obs = collections.OrderedDict()
    obs['egocentric_state'] = physics.egocentric_state()
    obs['torso_velocity'] = physics.torso_velocity()
    obs['torso_upright'] = physics.torso_upright()
    obs['imu'] = physics.imu()
    obs['force_torque'] = physics.force_torque()
    return obs

  def get_reward(self, physics):
    move_reward = rewards.tolerance(physics.torso_velocity()[:, 0], bounds=(self._desired_speed, float('inf')),
                                    margin=self._desired_speed, value_at_margin=0.5, sigmoid='linear')
    return _upright_reward(physics) * move_reward


def _make(speed):
  def make(batch=1, seed=0, time_limit=_DEFAULT_TIME_LIMIT, **physics_kw):
    physics_kw.setdefault('outputs', OUTPUTS)
    physics = Physics(testing_models.load('quadruped'), batch=batch, **physics_kw)
    return control.BatchedEnvironment(physics, Move(desired_speed=speed, seed=seed), time_limit=time_limit,
                                      control_timestep=_CONTROL_TIMESTEP)
  return make


TASKS = dict(walk=_make(_WALK_SPEED), run=_make(_RUN_SPEED))
