"""`BatchedComposerEnvironment`: B lock-stepped copies of `composer.Environment` (composer/environment.py:412-465).

The reference regains control in Python between the physics steps of a control step:

    hooks.before_step(physics, action)                    # task.before_step -> walker.apply_action
    for i in range(n_sub_steps):
      hooks.before_substep(physics, action)
      physics.step()                                      # legacy ordering: mj_step2 ... mj_step1
      hooks.after_substep(physics)                        # walker.after_substep -> mj_subtreeVel (legacy_base.py:179-186)
    hooks.after_step(physics)                             # task.after_step: contact scan, end-effector heights
    reward, discount, termination, observation

Here the same order holds with every quantity batched and resident on the device. A task that does not override the
substep hooks gets all `n_sub_steps` physics steps in ONE engine call (their only reference-side content,
`mj_subtreeVel`, is part of the engine's trailing position/velocity stage, so `data.subtree_linvel`, the contact list
and `xpos` are current when `after_step` runs); a task that does override them is stepped one physics step at a time
with the hooks in between — the engine keeps the state resident and starts each step from the position stage the
previous one left behind (B200MJ_STEP_REUSE_POS), so this costs launches, not recomputation.

`initialize_episode_mjcf` (the reference recompiles the model there, composer/environment.py:378-383) becomes
`task.initialize_episode_geoms(physics, env_mask)`: per-environment geom tables instead of a new model.
"""
from __future__ import annotations

import torch

from ..control import FIRST, LAST, MID, TimeStep, compute_n_steps


class Task:
  """Batched `composer.Task` surface (composer/task.py): the seven hooks + reward / discount / termination."""

  def initialize_episode_geoms(self, physics, env_mask):
    pass

  def initialize_episode(self, physics, env_mask):
    pass

  def before_step(self, physics, action):
    pass

  def before_substep(self, physics, action):
    pass

  def after_substep(self, physics):
    pass

  def after_step(self, physics):
    pass

  def get_reward(self, physics):
    raise NotImplementedError

  def get_discount(self, physics):
    return torch.ones(physics.batch, dtype=torch.float64, device=physics.device)

  def should_terminate_episode(self, physics):
    return torch.zeros(physics.batch, dtype=torch.bool, device=physics.device)

  def get_observation(self, physics):
    raise NotImplementedError


class BatchedComposerEnvironment:

  def __init__(self, physics, task, time_limit=float('inf'), physics_timestep=None, control_timestep=None, auto_reset=True):
    self._physics, self._task = physics, task
    physics.legacy_step = True            # composer/environment.py:173,309
    if physics_timestep is not None and abs(physics.timestep() - physics_timestep) > 1e-12:
      raise ValueError(f'model timestep {physics.timestep()} != task physics_timestep {physics_timestep}')
    self._n_sub_steps = compute_n_steps(control_timestep, physics.timestep()) if control_timestep else 1
    self._step_limit = float('inf') if time_limit == float('inf') else time_limit / (physics.timestep() * self._n_sub_steps)
    B, dev = physics.batch, physics.device
    self._step_count = torch.zeros(B, dtype=torch.int64, device=dev)
    self._reset_next = torch.ones(B, dtype=torch.bool, device=dev)
    self._auto_reset = auto_reset
    base = Task
    self._substep_hooks = (type(task).before_substep is not base.before_substep or type(task).after_substep is not base.after_substep)

  @property
  def physics(self):
    return self._physics

  @property
  def task(self):
    return self._task

  @property
  def n_sub_steps(self):
    return self._n_sub_steps

  def control_timestep(self):
    return self._physics.timestep() * self._n_sub_steps

  def _reset_envs(self, mask):
    # composer/environment.py:370-410: (re)build the arena for the episode, reset the physics, place the walker
    self._task.initialize_episode_geoms(self._physics, mask)
    self._physics.reset(env_mask=mask)
    self._task.initialize_episode(self._physics, mask)
    self._physics.after_reset(mask)

  def reset(self):
    self._reset_envs(None)
    self._step_count.zero_()
    self._reset_next.zero_()
    B = self._physics.batch
    return TimeStep(torch.full((B,), FIRST, device=self._physics.device), None, None, self._task.get_observation(self._physics))

  def step(self, action, timing=None):
    phys, task = self._physics, self._task
    if self._auto_reset and bool(self._reset_next.any()):
      mask = self._reset_next
      self._reset_envs(mask)
      self._step_count[mask] = 0
      self._reset_next = torch.zeros_like(mask)
    task.before_step(phys, action)
    if timing is not None:
      timing[0].record()
    if self._substep_hooks:
      for _ in range(self._n_sub_steps):
        task.before_substep(phys, action)
        phys.step(1)
        task.after_substep(phys)
    else:
      phys.step(self._n_sub_steps)
    if timing is not None:
      timing[1].record()
    task.after_step(phys)
    reward = task.get_reward(phys)
    discount = task.get_discount(phys)
    terminate = task.should_terminate_episode(phys)
    obs = task.get_observation(phys)
    self._step_count += 1
    last = terminate | (self._step_count >= self._step_limit)
    # the time limit ends an episode with discount 1, a task termination with the task's discount (environment.py:442-458)
    discount = torch.where(terminate, discount, torch.ones_like(discount))
    self._reset_next = last
    return TimeStep(torch.where(last, LAST, MID), reward, discount, obs)
