"""Batched LQR domain (reference: dm_control/suite/lqr.py:38-267). The Riccati solver the reference pairs with it
(suite/lqr_solver.py:27-82) is a test pin and lives in tests/lqr_riccati.py.

A chain of `n_bodies` unit spheres on collinear slide joints with random spring stiffness, the first `n_actuators`
of them driven by motors, constraints disabled: a linear system whose optimal policy and cost-to-go are known in
closed form. The reference uses it as a test of its physics (`suite/lqr_test.py:33-59`): the cost accumulated under
`u = K x` must equal `x0' P x0 / 2` to 1e-3. That test pins the joint-space inertia of a serial chain, the passive
spring force and the semi-implicit Euler update, independently of any MuJoCo binary; `tests/` runs it against the
oracle (CPU) and against the CUDA path.

One model is shared by the whole batch (as one reference environment owns one model); the initial states differ.
The spatial tendons the reference adds between consecutive bodies are for visualisation only (no stiffness, damping
or actuator refers to them) and are left out.
"""
from __future__ import annotations

import collections
import math

import numpy as np
import torch

from .. import control, mjcf_compile
from ..physics import BatchedPhysics
from . import base

_DEFAULT_TIME_LIMIT = float('inf')
_CONTROL_COST_COEF = 0.1
OUTPUTS = ('qM',)


def make_model_xml(n_bodies, n_actuators, random, stiffness_range=(15, 25), damping_range=(0, 0)):
  """MJCF of the spring chain (reference: lqr.py:137-196, `_make_body` :104-134; same draw order from `random`)."""
  if n_bodies < 1 or n_actuators < 1:
    raise ValueError('At least 1 body and 1 actuator required.')
  if n_actuators > n_bodies:
    raise ValueError('At most 1 actuator per body.')
  open_tags, motors = [], []
  for b in range(n_bodies):
    stiffness = random.uniform(stiffness_range[0], stiffness_range[1])
    damping = random.uniform(damping_range[0], damping_range[1])
    pos = '.25 0 .1' if b == 0 else '.25 0 0'
    open_tags.append(f'<body name="body_{b}" pos="{pos}"><joint name="joint_{b}" stiffness="{stiffness!r}" '
                     f'damping="{damping!r}"/><geom name="geom_{b}"/><site name="site_{b}"/>')
    if b < n_actuators:
      motors.append(f'<motor name="motor_{b}" joint="joint_{b}"/>')
  bodies = ''.join(open_tags) + '</body>' * n_bodies
  return f"""<mujoco model="LQR">
  <option timestep=".03"><flag constraint="disable"/></option>
  <default>
    <joint type="slide" axis="0 1 0"/>
    <geom type="sphere" size=".1"/>
    <site size=".01"/>
  </default>
  <worldbody>
    <geom name="floor" size="4 1 .2" type="plane"/>
    <geom name="origin" pos="2 0 .05" size="2 .003 .05" type="box"/>
    {bodies}
  </worldbody>
  <actuator>{''.join(motors)}</actuator>
</mujoco>"""


class Physics(BatchedPhysics):

  def state_norm(self):
    """[B] norm of (qpos, qvel) (reference: lqr.py:202-204)."""
    return torch.cat([self.data.qpos, self.data.qvel], dim=1).norm(dim=1)


class LQRLevel(base.Task):
  """cost = sum(position^2)/2 + c sum(control^2)/2 (reference: lqr.py:207-267)."""
  _TERMINAL_TOL = 1e-6

  def __init__(self, control_cost_coef, seed=0):
    if control_cost_coef <= 0:
      raise ValueError('control_cost_coef must be positive.')
    self._control_cost_coef = control_cost_coef
    super().__init__(seed)

  @property
  def control_cost_coef(self):
    return self._control_cost_coef

  def initialize_episode(self, physics, env_mask):
    """Random position on the sphere of radius sqrt(2), zero velocity."""
    gen = self.generator(physics)
    physics.reset(env_mask=env_mask)
    unit = torch.randn(physics.batch, physics.model.nq, generator=gen, device=physics.device, dtype=torch.float64)
    q = math.sqrt(2) * unit / unit.norm(dim=1, keepdim=True)
    if env_mask is None:
      physics.data.qpos.copy_(q)
    else:
      physics.data.qpos[env_mask] = q[env_mask]
    physics.after_reset()

  def get_observation(self, physics):
    obs = collections.OrderedDict()
    obs['position'] = physics.position().clone()
    obs['velocity'] = physics.velocity().clone()
    return obs

  def get_reward(self, physics):
    position, u = physics.position(), physics.control()
    return 1 - (0.5 * (position * position).sum(dim=1) + 0.5 * (u * u).sum(dim=1) * self._control_cost_coef)

  def get_evaluation(self, physics):
    return (physics.state_norm() <= 0.01).to(torch.float64)

  def get_termination(self, physics):
    """Discount 0 where the state norm fell below 1e-6, NaN (= keep going) elsewhere."""
    norm = physics.state_norm()
    return torch.where(norm < self._TERMINAL_TOL, torch.zeros_like(norm), torch.full_like(norm, float('nan')))


def compile_model(n_bodies, n_actuators, random):
  if not isinstance(random, np.random.RandomState):
    random = np.random.RandomState(random)
  return mjcf_compile.compile_xml(make_model_xml(n_bodies, n_actuators, random))


def _make(n_bodies, n_actuators):
  def make(batch=1, seed=0, time_limit=_DEFAULT_TIME_LIMIT, random=None, **physics_kw):
    physics_kw.setdefault('outputs', OUTPUTS)
    physics_kw.setdefault('sensors', False)
    physics = Physics(compile_model(n_bodies, n_actuators, seed if random is None else random), batch=batch, **physics_kw)
    return control.BatchedEnvironment(physics, LQRLevel(_CONTROL_COST_COEF, seed=seed), time_limit=time_limit)
  return make


TASKS = dict(lqr_2_1=_make(2, 1), lqr_6_2=_make(6, 2))
