"""`physics.bind(kind, names)`: the batched counterpart of `mjcf.Physics.bind` (dm_control/mjcf/physics.py:209-237,516-652).

The reference binds `mjcf.Element`s; the PyMJCF object model is out of scope here, so elements are named by
`(kind, names)` — `physics.bind('body', ['rradius', 'lradius']).xpos`, `physics.bind('joint', 'rfemurrx').qpos = 0.3`,
`physics.bind('geom', 'wall_3').pos`. What carries over is the contract:

  * one attribute namespace per element kind with the type prefix removed (`geom_pos` -> `.pos`, `body_mass` -> `.mass`,
    `jnt_range` -> `.range`, `actuator_gear` -> `.gear`), data fields by their own name (`xpos`, `qpos`, `qvel`, `ctrl`,
    `sensordata`, `subtree_linvel`, ...), `element_id`;
  * ragged fields resolve through the address tables (`qpos`/`qvel` of a joint, `sensordata` of a sensor);
  * writes to the STATE through a binding (`qpos qvel act ctrl qacc_warmstart time`) mark the physics dirty, and the
    next read of a DERIVED field through any binding runs `forward()` first — lazily, once
    (mjcf/physics.py:211-237 `_triggers_dirty` / `is_dirty`). Model writes re-upload the model before the next call.

Every value has the leading batch axis: `bind('body', names).xpos` is `[B, len(names), 3]` (or `[B, 3]` for one name).
"""
from __future__ import annotations

import numpy as np
import torch

_STATE = ('qpos', 'qvel', 'act', 'ctrl', 'qacc_warmstart', 'time')
_PREFIX = dict(body='body_', geom='geom_', site='site_', joint='jnt_', actuator='actuator_', sensor='sensor_', tendon='tendon_',
               dof='dof_')
# data fields addressable per element kind: name -> (row kind, ragged address table or None)
_DATA = {
    'body': ('xpos', 'xquat', 'xmat', 'xipos', 'subtree_com', 'subtree_linvel', 'cvel', 'xfrc_applied'),
    'geom': ('geom_xpos', 'geom_xmat'),
    'site': ('site_xpos', 'site_xmat'),
    'actuator': ('ctrl', 'actuator_force'),
}
_DATA_ALIAS = {'geom': {'xpos': 'geom_xpos', 'xmat': 'geom_xmat'}, 'site': {'xpos': 'site_xpos', 'xmat': 'site_xmat'},
               'actuator': {'force': 'actuator_force'}}


class Binding:

  def __init__(self, physics, kind, names):
    object.__setattr__(self, '_p', physics)
    object.__setattr__(self, '_kind', kind)
    single = isinstance(names, str)
    names = [names] if single else list(names)
    m = physics.model
    ids = np.array([m.name2id(n, kind) for n in names], dtype=np.int64)
    object.__setattr__(self, '_single', single)
    object.__setattr__(self, 'element_id', int(ids[0]) if single else ids)
    object.__setattr__(self, '_ids', ids)
    object.__setattr__(self, '_tids', torch.as_tensor(ids, device=physics.device))

  # ---- helpers ----
  def _ragged(self, field):
    """Flat column indices of a ragged data field for the bound elements (or None)."""
    m, ids = self._p.model, self._ids
    if self._kind == 'joint' and field in ('qpos', 'qvel', 'qacc', 'qacc_warmstart', 'qfrc_bias', 'qfrc_passive', 'qfrc_actuator', 'qfrc_constraint',
                                           'qfrc_applied'):
      qp = field == 'qpos'
      adr = np.asarray(m.jnt_qposadr if qp else m.jnt_dofadr)
      width = {0: (7, 6), 1: (4, 3), 2: (1, 1), 3: (1, 1)}
      cols = [a + k for j in ids for a in [int(adr[j])] for k in range(width[int(m.jnt_type[j])][0 if qp else 1])]
      return np.array(cols, dtype=np.int64)
    if self._kind == 'sensor' and field == 'sensordata':
      adr, dim = np.asarray(m.sensor_adr), np.asarray(m.sensor_dim)
      return np.array([int(adr[s]) + k for s in ids for k in range(int(dim[s]))], dtype=np.int64)
    if self._kind == 'actuator' and field == 'act':
      adr = np.asarray(m.actuator_actadr)
      return np.array([int(adr[a]) for a in ids if adr[a] >= 0], dtype=np.int64)
    return None

  def _data_field(self, name):
    d = self._p.data
    name = _DATA_ALIAS.get(self._kind, {}).get(name, name)
    t = getattr(d, name, None)
    return (name, t) if isinstance(t, torch.Tensor) else (name, None)

  def _squeeze(self, t):
    return t[:, 0] if self._single and t.dim() >= 2 else t

  # ---- attribute access ----
  def __getattr__(self, name):
    p = self._p
    field, t = self._data_field(name)
    if t is not None:
      if field not in _STATE and getattr(p, '_bind_dirty', False):
        p.forward()                              # lazily, once: derived quantities after a state write through a binding
      cols = self._ragged(field)
      if cols is not None:
        return t.index_select(1, torch.as_tensor(cols, device=p.device))
      if field in _DATA.get(self._kind, ()) or field in ('ctrl',):
        n = {'body': p.model.nbody, 'geom': p.model.ngeom, 'site': p.model.nsite, 'actuator': p.model.nu}[self._kind]
        return self._squeeze(t.reshape(p.batch, n, -1).index_select(1, self._tids)).squeeze(-1) if t.numel() // (p.batch * n) == 1 \
            else self._squeeze(t.reshape(p.batch, n, -1).index_select(1, self._tids))
      raise AttributeError(f'{name!r} is not a per-{self._kind} data field')
    mname = _PREFIX[self._kind] + name
    arr = p.model.fields.get(mname)
    if arr is None:
      raise AttributeError(f'bound {self._kind} has no attribute {name!r}')
    n = len(p.model.ordered_names[self._kind]) if self._kind in p.model.ordered_names else arr.shape[0]
    rows = np.asarray(arr).reshape(n, -1)[self._ids]
    rows = rows[:, 0] if rows.shape[1] == 1 else rows
    return rows[0] if self._single else rows

  def __setattr__(self, name, value):
    p = self._p
    field, t = self._data_field(name)
    if t is not None:
      v = torch.as_tensor(value, dtype=t.dtype, device=p.device)
      cols = self._ragged(field)
      if cols is not None:
        t[:, torch.as_tensor(cols, device=p.device)] = v
      else:
        n = {'body': p.model.nbody, 'geom': p.model.ngeom, 'site': p.model.nsite, 'actuator': p.model.nu}[self._kind]
        view = t.reshape(p.batch, n, -1)
        view[:, self._tids] = v.reshape((-1, len(self._ids), view.shape[2])) if v.numel() > view.shape[2] * len(self._ids) or v.dim() >= 2 \
            else v.reshape(1, -1, view.shape[2]).expand(p.batch, len(self._ids), view.shape[2])
      if field in _STATE:
        p.mark_as_dirty()
        if field != 'ctrl':                      # ctrl does not invalidate position / velocity-stage quantities
          object.__setattr__(p, '_bind_dirty', True)
      return
    mname = _PREFIX[self._kind] + name
    if mname not in p.model.fields:
      raise AttributeError(f'bound {self._kind} has no attribute {name!r}')
    arr = p.model.fields[mname]
    n = len(p.model.ordered_names[self._kind])
    arr.reshape(n, -1)[self._ids] = np.asarray(value, dtype=arr.dtype).reshape(len(self._ids), -1) if np.ndim(value) else value
    p.model.touch()
    object.__setattr__(p, '_bind_dirty', True)
