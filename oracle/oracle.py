"""ctypes front-end of the CPU oracle (oracle/mjoracle.cpp). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
PARITY UNPINNED — see the header of mjoracle.cpp and DESIGN.md.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libmjoracle.so')
_lib = None


def build(force=False):
  src = os.path.join(_HERE, 'mjoracle.cpp')
  inc = os.path.join(os.path.dirname(_HERE), 'include')
  deps = [src, os.path.join(inc, 'b200mj_model_fields.h'), os.path.join(inc, 'b200mj_convex.h')]
  if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
    subprocess.check_call(['make', '-C', _HERE, '-s'])
  return _SO


def lib():
  global _lib
  if _lib is None:
    build()   # make is a no-op when the .so is newer than its sources
    L = ctypes.CDLL(_SO)
    vp, ip, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    L.bmjo_model_create.restype = vp
    L.bmjo_model_create.argtypes = [ip, dp]
    L.bmjo_model_destroy.argtypes = [vp]
    L.bmjo_data_create.restype = vp
    L.bmjo_data_create.argtypes = [vp]
    L.bmjo_data_destroy.argtypes = [vp]
    L.bmjo_reset.argtypes = [vp, vp, ctypes.c_int]
    for f in ('bmjo_forward', 'bmjo_step1', 'bmjo_step2', 'bmjo_subtree_vel'):
      getattr(L, f).argtypes = [vp, vp]
    L.bmjo_step.argtypes = [vp, vp, ctypes.c_int]
    L.bmjo_control_step.argtypes = [vp, vp, ctypes.c_int]
    L.bmjo_rollout.argtypes = [vp, vp, dp, ctypes.c_int, ctypes.c_int]
    L.bmjo_set_disableflags.argtypes = [vp, ctypes.c_int]
    L.bmjo_get_disableflags.argtypes = [vp]
    L.bmjo_get_disableflags.restype = ctypes.c_int
    L.bmjo_field.restype = dp
    L.bmjo_field.argtypes = [vp, ctypes.c_char_p, ip]
    L.bmjo_efc_int.restype = ip
    L.bmjo_efc_int.argtypes = [vp, ctypes.c_char_p, ip]
    for f in ('bmjo_ncon', 'bmjo_nefc', 'bmjo_solver_niter'):
      getattr(L, f).argtypes = [vp]
      getattr(L, f).restype = ctypes.c_int
    L.bmjo_warning.restype = ip
    L.bmjo_warning.argtypes = [vp]
    L.bmjo_contact.argtypes = [vp, ctypes.c_int, dp]
    L.bmjo_narrowphase.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, dp, dp, dp, ctypes.c_double, dp]
    L.bmjo_narrowphase.restype = ctypes.c_int
    _lib = L
  return _lib


_SHAPES = dict(xpos=3, xquat=4, xmat=9, xipos=3, ximat=9, xanchor=3, xaxis=3, geom_xpos=3, geom_xmat=9, site_xpos=3,
               site_xmat=9, subtree_com=3, cinert=10, crb=10, cdof=6, cdof_dot=6, cvel=6, cacc=6, cfrc_int=6,
               cfrc_ext=6, subtree_linvel=3, xfrc_applied=6)


class Contact:
  __slots__ = ('dist', 'pos', 'frame', 'includemargin', 'friction', 'solref', 'solimp', 'dim', 'geom1', 'geom2',
               'efc_address')

  def __init__(self, raw):
    self.dist = raw[0]
    self.pos = raw[1:4].copy()
    self.frame = raw[4:13].copy()
    self.includemargin = raw[13]
    self.friction = raw[14:19].copy()
    self.solref = raw[19:21].copy()
    self.solimp = raw[21:26].copy()
    self.dim, self.geom1, self.geom2, self.efc_address = (int(x) for x in raw[26:30])


class OraclePhysics:
  """One environment stepped by the scalar oracle. Field access returns live numpy views."""

  def __init__(self, model):
    self.model = model
    self._L = lib()
    idata, rdata = model.pack()
    self._idata, self._rdata = idata, rdata
    self._m = self._L.bmjo_model_create(idata.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                        rdata.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    self._d = self._L.bmjo_data_create(self._m)
    self._views = {}

  def __del__(self):
    try:
      self._L.bmjo_data_destroy(self._d)
      self._L.bmjo_model_destroy(self._m)
    except Exception:
      pass

  def field(self, name):
    if name not in self._views:
      n = ctypes.c_int(0)
      p = self._L.bmjo_field(self._d, name.encode(), ctypes.byref(n))
      if n.value < 0:
        raise AttributeError(name)
      arr = np.ctypeslib.as_array(p, shape=(n.value,)) if n.value > 0 else np.zeros(0)
      w = _SHAPES.get(name)
      if w and n.value:
        arr = arr.reshape(-1, w)
      self._views[name] = arr
    return self._views[name]

  def __getattr__(self, name):
    if name.startswith('_'):
      raise AttributeError(name)
    return self.field(name)

  @property
  def time(self):
    return float(self.field('time')[0])

  @time.setter
  def time(self, v):
    self.field('time')[0] = v

  @property
  def ncon(self):
    return self._L.bmjo_ncon(self._d)

  @property
  def nefc(self):
    return self._L.bmjo_nefc(self._d)

  @property
  def solver_niter(self):
    return self._L.bmjo_solver_niter(self._d)

  @property
  def warning(self):
    return np.ctypeslib.as_array(self._L.bmjo_warning(self._d), shape=(8,))

  @property
  def contact(self):
    out = []
    buf = np.zeros(30)
    for i in range(self.ncon):
      self._L.bmjo_contact(self._d, i, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
      out.append(Contact(buf))
    return out

  def efc(self, name):
    n = self.nefc
    if name in ('efc_type', 'efc_id', 'efc_state'):
      k = ctypes.c_int(0)
      p = self._L.bmjo_efc_int(self._d, name.encode(), ctypes.byref(k))
      return np.ctypeslib.as_array(p, shape=(k.value,))[:n].copy()
    arr = self.field(name)
    if name == 'efc_J':
      return arr[:n * self.model.nv].reshape(n, self.model.nv).copy()
    return arr[:n].copy()

  def M_dense(self):
    nv = self.model.nv
    return self.field('M').reshape(nv, nv)

  def reset(self, key=-1):
    self._L.bmjo_reset(self._m, self._d, key)

  def forward(self):
    self._L.bmjo_forward(self._m, self._d)

  def step(self, n=1):
    self._L.bmjo_step(self._m, self._d, n)

  def step1(self):
    self._L.bmjo_step1(self._m, self._d)

  def step2(self):
    self._L.bmjo_step2(self._m, self._d)

  def control_step(self, nstep):
    """Reference legacy ordering: dm_control/mujoco/engine.py:147-162."""
    self._L.bmjo_control_step(self._m, self._d, nstep)

  def rollout(self, tape, nsub):
    """`len(tape)` control steps with actions tape[k] (C loop; the GIL is released for the whole call)."""
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    self._L.bmjo_rollout(self._m, self._d, tape.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), tape.shape[0], nsub)

  def subtree_vel(self):
    self._L.bmjo_subtree_vel(self._m, self._d)

  @property
  def disableflags(self):
    return self._L.bmjo_get_disableflags(self._m)

  @disableflags.setter
  def disableflags(self, v):
    self._L.bmjo_set_disableflags(self._m, int(v))

  def contact_force(self, i):
    """mj_contactForce equivalent: 6-vector (force in contact frame, torque) for contact i."""
    c = self.contact[i]
    out = np.zeros(6)
    if c.efc_address < 0:
      return out
    f = self.field('efc_force')
    if c.dim == 1:
      out[0] = f[c.efc_address]
    else:
      for k in range(1, c.dim):
        fp, fn = f[c.efc_address + 2 * (k - 1)], f[c.efc_address + 2 * (k - 1) + 1]
        out[0] += fp + fn
        out[k] = (fp - fn) * c.friction[k - 1]
    return out


def narrowphase(t1, p1, m1, s1, t2, p2, m2, s2, margin=0.0):
  """One geom pair through the oracle's narrow phase (types must satisfy t1 <= t2). Returns [(dist, pos, normal)]."""
  dp = ctypes.POINTER(ctypes.c_double)
  arrs = [np.ascontiguousarray(a, dtype=np.float64).ravel() for a in (p1, m1, s1, p2, m2, s2)]
  out = np.zeros(80)
  n = lib().bmjo_narrowphase(int(t1), arrs[0].ctypes.data_as(dp), arrs[1].ctypes.data_as(dp), arrs[2].ctypes.data_as(dp), int(t2),
                             arrs[3].ctypes.data_as(dp), arrs[4].ctypes.data_as(dp), arrs[5].ctypes.data_as(dp), float(margin),
                             out.ctypes.data_as(dp))
  return [(out[10 * k], out[10 * k + 1:10 * k + 4].copy(), out[10 * k + 4:10 * k + 7].copy()) for k in range(max(n, 0))]
