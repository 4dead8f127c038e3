"""`BatchedEnvironment`: B lock-stepped copies of `dm_control.rl.control.Environment` (rl/control.py:35-153).

Same call order per step (`before_step` -> `physics.step(n_sub_steps)` -> `after_step` -> reward -> observation),
same sub-step arithmetic (`compute_n_steps`, control.py:168-194), same time-limit rule, but every quantity has a
leading batch axis and stays on the device. Episodes that hit the time limit (or whose physics diverged) are reset
in place on the next `step`, per environment (the reference resets the single env on the next call, control.py:102).
"""
from __future__ import annotations

import collections

import torch


def compute_n_steps(control_timestep, physics_timestep, tolerance=1e-8):
  """Reference: rl/control.py:168-194 (same error text)."""
  if control_timestep < physics_timestep:
    raise ValueError('Control timestep ({}) cannot be smaller than physics timestep ({}).'.format(
        control_timestep, physics_timestep))
  if abs((control_timestep / physics_timestep - round(control_timestep / physics_timestep))) > tolerance:
    raise ValueError('Control timestep ({}) must be an integer multiple of physics timestep ({})'.format(
        control_timestep, physics_timestep))
  return int(round(control_timestep / physics_timestep))


TimeStep = collections.namedtuple('TimeStep', ['step_type', 'reward', 'discount', 'observation'])
FIRST, MID, LAST = 0, 1, 2


class BatchedEnvironment:

  def __init__(self, physics, task, time_limit=float('inf'), control_timestep=None, n_sub_steps=None,
               legacy_step=True, auto_reset=True, graph_task_ops=False):
    self._physics, self._task = physics, task
    physics.legacy_step = legacy_step
    if n_sub_steps is not None and control_timestep is not None:
      raise ValueError('Both n_sub_steps and control_timestep were supplied.')
    if n_sub_steps is not None:
      self._n_sub_steps = n_sub_steps
    elif control_timestep is not None:
      self._n_sub_steps = compute_n_steps(control_timestep, physics.timestep())
    else:
      self._n_sub_steps = 1
    self._step_limit = float('inf') if time_limit == float('inf') else time_limit / (physics.timestep() * self._n_sub_steps)
    self._step_count = torch.zeros(physics.batch, dtype=torch.int64, device=physics.device)
    self._reset_next = torch.ones(physics.batch, dtype=torch.bool, device=physics.device)
    self._auto_reset = auto_reset
    # host-side upper bound of every environment's step count: while it is below the step limit no environment can
    # be LAST, so step() need not read `_reset_next` back (a device->host sync per step otherwise)
    self._count_ub = float('inf')     # unknown until the first check
    # Optional: replay the task's reward/observation torch ops (~80 tiny launches) as one CUDA graph. The returned
    # reward / observation tensors are then static buffers that the next step overwrites.
    self._graph_task_ops = graph_task_ops
    self._graph = None
    self._graph_out = None

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

  def reset(self):
    self._task.initialize_episode(self._physics, None)
    self._step_count.zero_()
    self._reset_next.zero_()
    self._count_ub = 0
    obs = self._task.get_observation(self._physics)
    B = self._physics.batch
    return TimeStep(torch.full((B,), FIRST, device=self._physics.device), None, None, obs)

  def _reward_and_observation(self):
    if not self._graph_task_ops:
      return self._task.get_reward(self._physics), self._task.get_observation(self._physics)
    if self._graph is None:
      # warm up eagerly on a side stream (allocator + lazy module loads), then capture once
      s = torch.cuda.Stream(device=self._physics.device)
      s.wait_stream(torch.cuda.current_stream(self._physics.device))
      with torch.cuda.stream(s):
        for _ in range(2):
          self._task.get_reward(self._physics); self._task.get_observation(self._physics)
      torch.cuda.current_stream(self._physics.device).wait_stream(s)
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        out = (self._task.get_reward(self._physics), self._task.get_observation(self._physics))
      self._graph, self._graph_out = g, out
    self._graph.replay()
    return self._graph_out

  def step(self, action, timing=None):
    """One control step for every environment. `timing`: optional (start, end) CUDA events recorded around the
    physics call (bench.py times the step's kernel group with them); no effect on the result."""
    if self._auto_reset and self._count_ub >= self._step_limit:
      if bool(self._reset_next.any()):
        mask = self._reset_next
        self._task.initialize_episode(self._physics, mask)
        self._step_count[mask] = 0
        self._reset_next = torch.zeros_like(mask)
      self._count_ub = int(self._step_count.max())      # slow path only: one more readback, then exact again
    self._task.before_step(action, self._physics)
    if timing is not None:
      timing[0].record()
    self._physics.step(self._n_sub_steps)
    if timing is not None:
      timing[1].record()
    self._task.after_step(self._physics)
    reward, obs = self._reward_and_observation()
    self._step_count += 1
    self._count_ub += 1
    last = self._step_count >= self._step_limit
    discount = torch.ones_like(reward)
    # task termination (control.py:113-118): None = the task never terminates; otherwise a [B] tensor holding the
    # terminal discount for the environments that end now and NaN for the others. The time limit wins (discount 1).
    get_term = getattr(self._task, 'get_termination', None)
    term = get_term(self._physics) if get_term is not None else None
    if term is not None:
      ended = ~torch.isnan(term)
      discount = torch.where(ended & ~last, term, discount)
      last = last | ended
      self._count_ub = float('inf')       # episodes may end at any step: check the flags on the next call
    self._reset_next = last
    step_type = torch.where(last, LAST, MID)
    return TimeStep(step_type, reward, discount, obs)
