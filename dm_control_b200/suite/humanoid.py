"""Batched humanoid stand/walk/run (reference: dm_control/suite/humanoid.py)."""
from __future__ import annotations

import collections

import torch

from .. import control, rewards, testing_models
from ..physics import BatchedPhysics
from . import base

_DEFAULT_TIME_LIMIT = 25
_CONTROL_TIMESTEP = .025
_STAND_HEIGHT = 1.4
_WALK_SPEED = 1
_RUN_SPEED = 10

OUTPUTS = ('xpos', 'xmat', 'subtree_com', 'sensordata', 'ncon', 'nefc', 'solver_niter')


class Physics(BatchedPhysics):
  """Accessors of the reference `humanoid.Physics` (humanoid.py:93-129), batched."""

  def _ids(self):
    if not hasattr(self, '_hid'):
      n = self.model.names
      limbs = [n['body'][s + l] for s in ('left_', 'right_') for l in ('hand', 'foot')]
      self._hid = dict(torso=n['body']['torso'], head=n['body']['head'],
                       limbs=torch.tensor(limbs, dtype=torch.int64, device=self.device),   # device-resident: no per-step H2D
                       linvel=int(self.model.sensor_adr[n['sensor']['torso_subtreelinvel']]))
    return self._hid

  def torso_upright(self):
    return self.data.xmat[:, self._ids()['torso'], 8]

  def head_height(self):
    return self.data.xpos[:, self._ids()['head'], 2]

  def center_of_mass_position(self):
    return self.data.subtree_com[:, self._ids()['torso']].clone()

  def center_of_mass_velocity(self):
    a = self._ids()['linvel']
    return self.data.sensordata[:, a:a + 3].clone()

  def torso_vertical_orientation(self):
    return self.data.xmat[:, self._ids()['torso'], 6:9]

  def joint_angles(self):
    return self.data.qpos[:, 7:].clone()

  def extremities(self):
    ids = self._ids()
    frame = self.data.xmat[:, ids['torso']].reshape(-1, 3, 3)
    torso_pos = self.data.xpos[:, ids['torso']]
    rel = self.data.xpos.index_select(1, ids['limbs']) - torso_pos[:, None, :]   # [B, 4, 3]
    return torch.bmm(rel, frame).reshape(self.batch, 12)                   # row-vector . frame (humanoid.py:120-129)


class Humanoid(base.Task):

  def __init__(self, move_speed, pure_state, seed=0):
    self._move_speed, self._pure_state = move_speed, pure_state
    super().__init__(seed)

  def initialize_episode(self, physics, env_mask):
    """Reject-sample collision-free configurations per environment (humanoid.py:152-166)."""
    gen = self.generator(physics)
    todo = torch.ones(physics.batch, dtype=torch.bool, device=physics.device) if env_mask is None else env_mask.clone()
    physics.reset(env_mask=None if env_mask is None else env_mask)
    for _ in range(200):
      base.randomize_limited_and_rotational_joints(physics, gen, todo)
      physics.after_reset(todo)      # only the environments still being drawn: the others keep their outputs
      todo = todo & (physics.data.ncon > 0)
      if not bool(todo.any()):
        break
    else:
      # the reference loops until the pose is contact-free (humanoid.py:160-166); a batch cannot loop forever
      raise RuntimeError(f'humanoid reset: {int(todo.sum())} environments still in contact after 200 re-draws')

  def get_observation(self, physics):
    obs = collections.OrderedDict()
    if self._pure_state:
      obs['position'] = physics.position().clone()
      obs['velocity'] = physics.velocity().clone()
    else:
      obs['joint_angles'] = physics.joint_angles()
      obs['head_height'] = physics.head_height()
      obs['extremities'] = physics.extremities()
      obs['torso_vertical'] = physics.torso_vertical_orientation()
      obs['com_velocity'] = physics.center_of_mass_velocity()
      obs['velocity'] = physics.velocity().clone()
    return obs

  def get_reward(self, physics):
    standing = rewards.tolerance(physics.head_height(), bounds=(_STAND_HEIGHT, float('inf')), margin=_STAND_HEIGHT / 4)
    upright = rewards.tolerance(physics.torso_upright(), bounds=(0.9, float('inf')), sigmoid='linear', margin=1.9,
                                value_at_margin=0)
    stand_reward = standing * upright
    small_control = rewards.tolerance(physics.control(), margin=1, value_at_margin=0, sigmoid='quadratic').mean(dim=1)
    small_control = (4 + small_control) / 5
    if self._move_speed == 0:
      horizontal_velocity = physics.center_of_mass_velocity()[:, 0:2]
      dont_move = rewards.tolerance(horizontal_velocity, margin=2).mean(dim=1)
      return small_control * stand_reward * dont_move
    com_velocity = physics.center_of_mass_velocity()[:, 0:2].norm(dim=1)
    move = rewards.tolerance(com_velocity, bounds=(self._move_speed, float('inf')), margin=self._move_speed,
                             value_at_margin=0, sigmoid='linear')
    move = (5 * move + 1) / 6
    return small_control * stand_reward * move


def _make(move_speed, pure_state=False):
  def make(batch=1, seed=0, time_limit=_DEFAULT_TIME_LIMIT, **physics_kw):
    physics_kw.setdefault('outputs', OUTPUTS)
    # the trailing mj_step1 of a step runs collision + constraint assembly like the reference's (data.ncon and the
    # contact list are those of the new state); its results are what the next step() starts from
    # (B200MJ_STEP_REUSE_POS), so this costs no extra position stage
    physics = Physics(testing_models.load('humanoid'), batch=batch, **physics_kw)
    task = Humanoid(move_speed=move_speed, pure_state=pure_state, seed=seed)
    return control.BatchedEnvironment(physics, task, time_limit=time_limit, control_timestep=_CONTROL_TIMESTEP)
  return make


TASKS = dict(stand=_make(0), walk=_make(_WALK_SPEED), run=_make(_RUN_SPEED), run_pure_state=_make(_RUN_SPEED, True))
