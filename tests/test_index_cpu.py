"""Named indexing twin (dm_control/mujoco/index.py semantics) on batched tensors — no GPU needed."""
import types

import numpy as np
import pytest
import torch

from dm_control_b200 import index
from dm_control_b200 import testing_models as tm


def _fake_physics(name, B=3):
  m = tm.load(name)
  d = types.SimpleNamespace()
  d.xpos = torch.arange(B * m.nbody * 3, dtype=torch.float64).reshape(B, m.nbody, 3)
  d.xmat = torch.arange(B * m.nbody * 9, dtype=torch.float64).reshape(B, m.nbody, 9)
  d.qpos = torch.arange(B * m.nq, dtype=torch.float64).reshape(B, m.nq)
  d.qvel = torch.arange(B * m.nv, dtype=torch.float64).reshape(B, m.nv)
  d.ctrl = torch.zeros(B, m.nu, dtype=torch.float64)
  d.sensordata = torch.arange(B * m.nsensordata, dtype=torch.float64).reshape(B, m.nsensordata)
  return types.SimpleNamespace(model=m, data=d)


def test_rows_columns_and_batch_axis():
  p = _fake_physics('humanoid')
  n = index.NamedIndexStructs(p)
  m = p.model
  torso, head = m.name2id('torso', 'body'), m.name2id('head', 'body')
  assert torch.equal(n.data.xpos['head', 'z'], p.data.xpos[:, head, 2])                 # humanoid.py:100-102
  assert torch.equal(n.data.xmat['torso', 'zz'], p.data.xmat[:, torso, 8])               # humanoid.py:96-98
  assert torch.equal(n.data.xmat['torso', ['zx', 'zy', 'zz']], p.data.xmat[:, torso, 6:9])
  assert torch.equal(n.data.xpos[['left_hand', 'right_foot']],
                     p.data.xpos[:, [m.name2id('left_hand', 'body'), m.name2id('right_foot', 'body')]])
  assert n.data.xpos[['torso', 'head'], ['x', 'z']].shape == (3, 2, 2)                   # outer product of names
  assert torch.equal(n.data.xpos[2], p.data.xpos[:, 2])                                   # plain ints still work


def test_ragged_axes():
  p = _fake_physics('humanoid')
  n = index.NamedIndexStructs(p)
  assert n.data.qpos['root'].shape == (3, 7)                       # free joint: 7 qpos, 6 qvel (index.py:94-100)
  assert n.data.qvel['root'].shape == (3, 6)
  assert torch.equal(n.data.qpos['abdomen_z'], p.data.qpos[:, 7:8])
  a = int(p.model.sensor_adr[p.model.names['sensor']['torso_subtreelinvel']])
  assert torch.equal(n.data.sensordata['torso_subtreelinvel'], p.data.sensordata[:, a:a + 3])
  n.data.qpos['abdomen_z'] = 0.25
  assert float(p.data.qpos[1, 7]) == 0.25
  n.data.ctrl['right_knee'] = torch.tensor([1.0, 2.0, 3.0])
  assert p.data.ctrl[:, p.model.name2id('right_knee', 'actuator')].tolist() == [1.0, 2.0, 3.0]


def test_model_side_and_errors():
  p = _fake_physics('cheetah')
  n = index.NamedIndexStructs(p)
  np.testing.assert_allclose(n.model.jnt_range['bthigh'], np.deg2rad([-30, 60]))
  assert n.model.geom_size['torso', 'x'] == pytest.approx(0.046)
  with pytest.raises(IndexError):
    n.data.xpos['no_such_body']
  with pytest.raises(IndexError):
    n.data.xpos['torso', 'w']
  with pytest.raises(AttributeError):
    n.data.not_a_field


def test_model_writes_are_tracked_or_refused():
  """Editing the model must reach the device copy: named writes and model.set() bump the version that triggers the
  re-upload; a raw `model.<field>[i] = v` is refused (read-only view) instead of being silently ignored."""
  import types
  import numpy as np
  import pytest
  from dm_control_b200 import index, testing_models as tm
  m = tm.load('cheetah').copy()
  phys = types.SimpleNamespace(model=m, data=types.SimpleNamespace())
  named = index.NamedIndexStructs(phys)
  v0 = m._version
  named.model.geom_size['ground', 0] = 3.5
  assert m._version == v0 + 1 and m.fields['geom_size'].reshape(-1, 3)[m.name2id('ground', 'geom'), 0] == 3.5
  m.set('body_mass', 2.0, 1)
  assert m._version == v0 + 2 and m.body_mass[1] == 2.0
  with pytest.raises(ValueError):
    m.body_mass[1] = 3.0
