"""Batched twin of `suite/utils/randomizers.py` against the reference's own test cases
(dm_control/suite/utils/randomizers_test.py:30-160), on host tensors — no GPU needed."""
import math
import types

import numpy as np
import torch

from dm_control_b200 import index, mjcf_compile
from dm_control_b200.suite import base


def _physics(xml, B=16):
  m = mjcf_compile.compile_xml(xml)
  d = types.SimpleNamespace(qpos=torch.as_tensor(np.asarray(m.qpos0)).repeat(B, 1).clone())
  p = types.SimpleNamespace(model=m, data=d, batch=B, device=torch.device('cpu'))
  p.named = index.NamedIndexStructs(p)
  return p


def test_single_joint_of_each_type():
  p = _physics("""<mujoco><default><joint range="0 90" armature="1"/></default><worldbody>
      <body><geom type="box" size="1 1 1"/><joint name="free" type="free"/></body>
      <body><geom type="box" size="1 1 1"/><joint name="limited_hinge" type="hinge" limited="true"/>
        <joint name="slide" type="slide" limited="false"/><joint name="limited_slide" type="slide" limited="true"/>
        <joint name="hinge" type="hinge" limited="false"/></body>
      <body><geom type="box" size="1 1 1"/><joint name="ball" type="ball" limited="false"/></body>
      <body><geom type="box" size="1 1 1"/><joint name="limited_ball" type="ball" limited="true"/></body>
    </worldbody></mujoco>""")
  base.randomize_limited_and_rotational_joints(p, torch.Generator().manual_seed(100))
  q = p.named.data.qpos
  for name in ('hinge', 'limited_hinge', 'limited_slide'):
    assert bool((q[name] != 0).all())
  for name in ('ball', 'limited_ball'):
    assert torch.allclose(q[name].norm(dim=1), torch.ones(16, dtype=torch.float64))
    assert bool((q[name][:, 1:].abs().sum(dim=1) > 0).all())
  assert torch.allclose(q['free'][:, 3:].norm(dim=1), torch.ones(16, dtype=torch.float64))
  # unlimited slide and the positional part of the free joint stay where they were
  assert bool((q['slide'] == 0).all()) and bool((q['free'][:, :3] == 0).all())
  # environments are randomised independently
  assert len(set(q['hinge'][:, 0].tolist())) == 16


def test_ranges_are_respected():
  p = _physics("""<mujoco><default><joint limited="true"/></default><worldbody><body><geom type="box" size="1 1 1"/>
      <joint name="hinge" type="hinge" range="0 10"/><joint name="slide" type="slide" range="30 50"/>
      <joint name="free_hinge" type="hinge" axis="0 1 0" limited="false"/></body>
      <body name="b" zaxis="1 0 0"><geom type="box" size="1 1 1"/><joint name="ball" type="ball" range="0 60"/></body>
    </worldbody></mujoco>""", B=64)
  g = torch.Generator().manual_seed(1)
  for _ in range(10):
    base.randomize_limited_and_rotational_joints(p, g)
    q = p.named.data.qpos
    assert float(q['hinge'].min()) >= 0 and float(q['hinge'].max()) <= math.radians(10)
    assert float(q['slide'].min()) >= 30 and float(q['slide'].max()) <= 50
    assert float(q['free_hinge'].abs().max()) <= math.pi
    # limited ball: the rotation angle stays inside the cone (randomizers_test.py:134-160)
    angle = 2 * torch.acos(q['ball'][:, 0].clamp(-1, 1))
    assert float(angle.max()) <= math.radians(60) + 1e-12 and float(angle.min()) >= 0


def test_masked_randomisation_leaves_other_environments_alone():
  p = _physics('<mujoco><worldbody><body><geom type="box" size="1 1 1"/><joint name="h" type="hinge"/></body></worldbody></mujoco>', B=8)
  mask = torch.tensor([True, False] * 4)
  base.randomize_limited_and_rotational_joints(p, torch.Generator().manual_seed(3), env_mask=mask)
  q = p.data.qpos[:, 0]
  assert bool((q[~mask] == 0).all()) and bool((q[mask] != 0).all())
