"""Batched `suite.base.Task` (reference: dm_control/suite/base.py:24-93)."""
from __future__ import annotations

import torch


class Task:

  def __init__(self, seed=0):
    self._seed = seed
    self._gen = None

  def generator(self, physics):
    if self._gen is None:
      self._gen = torch.Generator(device=physics.device).manual_seed(self._seed)
    return self._gen

  def initialize_episode(self, physics, env_mask):
    raise NotImplementedError

  def before_step(self, action, physics):
    """Reference: base.py:73-77 -> physics.set_control(action)."""
    physics.set_control(action)

  def after_step(self, physics):
    pass

  def get_termination(self, physics):
    """Reference: base.py (`get_termination` returns None = continue). Batched form: None, or a [B] tensor with the
    terminal discount where an episode ends now and NaN elsewhere."""
    return None

  def action_spec(self, physics):
    m = physics.model
    lo = torch.as_tensor(m.actuator_ctrlrange[:, 0].copy())
    hi = torch.as_tensor(m.actuator_ctrlrange[:, 1].copy())
    lim = torch.as_tensor(m.actuator_ctrllimited.astype(bool))
    big = torch.full_like(lo, 1e10)
    return torch.where(lim, lo, -big), torch.where(lim, hi, big)


def randomize_limited_and_rotational_joints(physics, gen, env_mask=None):
  """Batched `suite/utils/randomizers.py:35-88`: limited hinge/slide uniform in range, limited ball within its cone,
  unlimited hinge in [-pi, pi], unlimited ball uniform on the 3-sphere, free-joint quaternion from `rand(4)`
  normalised (the reference's documented quirk); slides without limits and free-joint positions stay put."""
  import math
  m, d = physics.model, physics.data
  B = physics.batch
  q = d.qpos.clone()
  for j in range(m.njnt):
    t, qa = int(m.jnt_type[j]), int(m.jnt_qposadr[j])
    lo, hi = float(m.jnt_range[j, 0]), float(m.jnt_range[j, 1])
    if m.jnt_limited[j]:
      if t in (2, 3):
        q[:, qa] = torch.rand(B, generator=gen, device=physics.device, dtype=torch.float64) * (hi - lo) + lo
      elif t == 1:
        # limited ball joint: random axis, rotation angle uniform in [0, range_max] (randomizers.py:25-33)
        axis = torch.randn(B, 3, generator=gen, device=physics.device, dtype=torch.float64)
        axis = axis / axis.norm(dim=1, keepdim=True)
        half = 0.5 * hi * torch.rand(B, 1, generator=gen, device=physics.device, dtype=torch.float64)
        q[:, qa:qa + 4] = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    else:
      if t == 3:
        q[:, qa] = (torch.rand(B, generator=gen, device=physics.device, dtype=torch.float64) * 2 - 1) * math.pi
      elif t == 1:
        quat = torch.randn(B, 4, generator=gen, device=physics.device, dtype=torch.float64)
        q[:, qa:qa + 4] = quat / quat.norm(dim=1, keepdim=True)
      elif t == 0:
        quat = torch.rand(B, 4, generator=gen, device=physics.device, dtype=torch.float64)
        q[:, qa + 3:qa + 7] = quat / quat.norm(dim=1, keepdim=True)
  if env_mask is None:
    d.qpos.copy_(q)
  else:
    d.qpos[env_mask] = q[env_mask]
