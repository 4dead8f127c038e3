"""Named indexing on batched tensors: `physics.named.data.xpos['torso', 'z']`, `named.data.qpos['root']`, ...

Batched twin of `dm_control/mujoco/index.py` (row names from the model's name tables, ragged `nq/nv/na/
nsensordata` axes via `jnt_qposadr/jnt_dofadr/actuator_actadr/sensor_adr` — index.py:94-100 — and the column names
`x y z`, `qw qx qy qz`, `xx … zz` — index.py:103-175). Every field keeps its leading batch axis: an indexing
expression that yields shape `S` in the reference yields `[B, *S]` here. Model-side fields (`named.model.*`) are
unbatched numpy arrays and index exactly as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

_COLS = {3: ['x', 'y', 'z'], 4: ['qw', 'qx', 'qy', 'qz'], 9: ['xx', 'xy', 'xz', 'yx', 'yy', 'yz', 'zx', 'zy', 'zz']}
_RGBA = {'geom_rgba', 'site_rgba', 'mat_rgba'}      # columns r g b a (index.py:103-175 `rgba`); visual tables, Model.vis
_XYZ = {'body_pos', 'body_ipos', 'body_inertia', 'jnt_pos', 'jnt_axis', 'geom_size', 'geom_pos', 'site_size', 'site_pos',
        'xpos', 'xipos', 'xanchor', 'xaxis', 'geom_xpos', 'site_xpos', 'subtree_com', 'subtree_linvel'}
_QUAT = {'body_quat', 'body_iquat', 'geom_quat', 'site_quat', 'xquat'}
_MAT = {'xmat', 'ximat', 'geom_xmat', 'site_xmat'}

# field -> kind of its row axis
_ROW_KIND = {
    **{f: 'body' for f in ('xpos', 'xquat', 'xmat', 'xipos', 'subtree_com', 'subtree_linvel', 'cvel', 'xfrc_applied',
                           'body_pos', 'body_quat', 'body_ipos', 'body_iquat', 'body_mass', 'body_inertia',
                           'body_parentid', 'body_rootid', 'body_subtreemass')},
    **{f: 'geom' for f in ('geom_xpos', 'geom_xmat', 'geom_size', 'geom_pos', 'geom_quat', 'geom_type', 'geom_bodyid',
                           'geom_friction', 'geom_rbound', 'geom_condim', 'geom_rgba', 'geom_group')},
    **{f: 'site' for f in ('site_xpos', 'site_xmat', 'site_pos', 'site_quat', 'site_size', 'site_bodyid', 'site_rgba', 'site_group')},
    **{f: 'camera' for f in ('cam_pos', 'cam_quat', 'cam_fovy', 'cam_bodyid', 'cam_mode', 'cam_targetbodyid')},
    'mat_rgba': 'material',
    **{f: 'joint' for f in ('jnt_type', 'jnt_range', 'jnt_limited', 'jnt_pos', 'jnt_axis', 'jnt_stiffness', 'jnt_qposadr',
                            'jnt_dofadr', 'jnt_bodyid')},
    **{f: 'actuator' for f in ('ctrl', 'actuator_force', 'actuator_ctrlrange', 'actuator_gear', 'actuator_ctrllimited')},
    **{f: 'nq' for f in ('qpos', 'qpos0')},
    **{f: 'nv' for f in ('qvel', 'qacc', 'qacc_warmstart', 'qfrc_bias', 'qfrc_passive', 'qfrc_actuator', 'qfrc_constraint',
                         'qfrc_applied', 'dof_damping', 'dof_armature')},
    'act': 'na', 'sensordata': 'nsensordata',
}


def _ragged(model, kind):
  """name -> slice for the ragged axes (index.py:245-267)."""
  out = {}
  if kind == 'nq':
    adr, names, total = model.jnt_qposadr, model.ordered_names['joint'], model.nq
  elif kind == 'nv':
    adr, names, total = model.jnt_dofadr, model.ordered_names['joint'], model.nv
  elif kind == 'nsensordata':
    adr, names, total = model.sensor_adr, model.ordered_names['sensor'], model.nsensordata
  else:   # 'na': singleton per actuator, -1 = none
    for i, n in enumerate(model.ordered_names['actuator']):
      a = int(model.actuator_actadr[i])
      if a >= 0:
        out[n] = slice(a, a + 1)
    return out
  ends = list(adr[1:]) + [total]
  for n, a, b in zip(names, adr, ends):
    out[n] = slice(int(a), int(b))
  return out


class FieldIndexer:
  """One field with name-aware `__getitem__` / `__setitem__` (index.py:455-600)."""

  def __init__(self, name, array, model, batched, on_write=None):
    self._name, self._a, self._batched, self._on_write = name, array, batched, on_write
    kind = _ROW_KIND.get(name)
    self._ragged = kind in ('nq', 'nv', 'na', 'nsensordata')
    if kind is None:
      self._rows = {}
    elif self._ragged:
      self._rows = _ragged(model, kind)
    else:
      self._rows = {n: i for i, n in enumerate(model.ordered_names.get(kind, []))}
    ncol = array.shape[-1] if array.ndim - (1 if batched else 0) >= 2 else 0
    self._cols = {c: i for i, c in enumerate(_COLS.get(ncol, []))} if (name in _XYZ or name in _QUAT or name in _MAT) else {}
    if name in _RGBA:
      self._cols = {'r': 0, 'g': 1, 'b': 2, 'a': 3}

  def _row_key(self, k):
    if isinstance(k, str):
      if k not in self._rows:
        raise IndexError(f'{k!r} is not a valid row name for field {self._name!r}; valid: {sorted(self._rows)[:8]}...')
      return self._rows[k]
    if isinstance(k, (list, tuple)) and k and all(isinstance(x, str) for x in k):
      idx = []
      for x in k:
        r = self._row_key(x)
        idx.extend(range(r.start, r.stop) if isinstance(r, slice) else [r])
      return idx
    return k

  def _col_key(self, k):
    if isinstance(k, str):
      if k not in self._cols:
        raise IndexError(f'{k!r} is not a valid column name for field {self._name!r}')
      return self._cols[k]
    if isinstance(k, (list, tuple)) and k and all(isinstance(x, str) for x in k):
      return [self._col_key(x) for x in k]
    return k

  def _translate(self, key):
    if not isinstance(key, tuple):
      key = (key,)
    out = [self._row_key(key[0])] + ([self._col_key(key[1])] if len(key) > 1 else []) + list(key[2:])
    if len(out) == 2 and isinstance(out[0], list) and isinstance(out[1], list):
      out = [np.asarray(out[0])[:, None], np.asarray(out[1])[None, :]]     # outer indexing, as np.ix_ in the reference
    return tuple([slice(None)] + out) if self._batched else tuple(out)

  def __getitem__(self, key):
    return self._a[self._translate(key)]

  def __setitem__(self, key, value):
    k = self._translate(key)
    if isinstance(self._a, torch.Tensor):
      self._a[k] = torch.as_tensor(value, dtype=self._a.dtype, device=self._a.device)
    else:
      self._a[k] = value
      if self._on_write is not None:
        self._on_write()          # model field: the device copy must be refreshed before the next step

  @property
  def row_names(self):
    return list(self._rows)

  def __repr__(self):
    return f'FieldIndexer({self._name}, rows={len(self._rows)}, shape={tuple(self._a.shape)})'


class _Struct:
  def __init__(self, getter, model, batched):
    object.__setattr__(self, '_g', getter)
    object.__setattr__(self, '_m', model)
    object.__setattr__(self, '_b', batched)
    object.__setattr__(self, '_cache', {})

  def __getattr__(self, name):
    arr = self._g(name)
    if arr is None:
      raise AttributeError(name)
    c = self._cache.get(name)
    if c is None or c._a is not arr:
      c = FieldIndexer(name, arr, self._m, self._b, on_write=None if self._b else self._m.touch)
      self._cache[name] = c
    return c


class NamedIndexStructs:
  """`physics.named` (engine.py:427-430): `.data.<field>[names]` on `[B, ...]` tensors, `.model.<field>[names]` on numpy."""

  def __init__(self, physics):
    m = physics.model

    def data_get(name):
      t = getattr(physics.data, name, None)
      if t is None:
        return None
      # [B, nbody*3]-style fields are exposed as [B, nbody, 3] (they are allocated that way already)
      return t

    def model_get(name):
      a = m.fields.get(name)
      if a is None:
        a = getattr(m, 'vis', {}).get(name)      # colours, groups, cameras: visual tables outside the physics blob
      if a is None:
        return None
      w = {'body_pos': 3, 'body_quat': 4, 'body_ipos': 3, 'body_iquat': 4, 'body_inertia': 3, 'geom_size': 3, 'geom_pos': 3,
           'geom_quat': 4, 'geom_friction': 3, 'site_pos': 3, 'site_quat': 4, 'site_size': 3, 'jnt_pos': 3, 'jnt_axis': 3,
           'jnt_range': 2, 'actuator_ctrlrange': 2}.get(name)
      return a.reshape(-1, w) if (w and a.ndim == 1) else a
    self.data = _Struct(data_get, m, True)
    self.model = _Struct(model_get, m, False)
