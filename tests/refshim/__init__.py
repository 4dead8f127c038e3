"""Import shims that let the UNMODIFIED reference sources under /root/reference/dm_control run on this engine.

`install()` registers, in sys.modules:
  * `dm_env` (+ `dm_env.specs`): a minimal stand-in for the DeepMind `dm_env` package (absent from this image) — the
    TimeStep / StepType / specs surface that `dm_control/rl/control.py` uses;
  * `dm_control`, `dm_control.rl`, `dm_control.suite`, `dm_control.suite.utils`, `dm_control.utils`: EMPTY package objects
    whose `__path__` points INTO /root/reference, so `import dm_control.rl.control`, `dm_control.suite.humanoid`,
    `dm_control.suite.base`, `dm_control.suite.common`, `dm_control.suite.utils.randomizers`, `dm_control.utils.rewards`
    ... execute the reference's own files, unmodified, without running the package `__init__`s that would pull in
    `mujoco`, `lxml`, OpenGL, ...;
  * `lxml.etree` (absent from this image): the few ElementTree calls suite/cartpole.py, suite/quadruped.py and
    utils/xml_tools.py make, on top of xml.etree.ElementTree;
  * `dm_control.mujoco`: a stand-in module whose `Physics` is `dm_control_b200.refview.SingleEnvPhysics` (the reference's
    numpy Physics API on a B = 1 view of the batched CUDA engine), plus `action_spec` and the `wrapper.mjbindings.enums`
    the randomizers read.
Test infrastructure only; nothing under dm_control_b200/ imports it."""
import collections
import enum
import os
import sys
import types

import numpy as np

REFERENCE = os.environ.get('DM_CONTROL_REFERENCE', '/root/reference')


def available():
  return os.path.isdir(os.path.join(REFERENCE, 'dm_control', 'suite'))


# ---------------------------------------------------------------------------------------------------------------
# dm_env stand-in
# ---------------------------------------------------------------------------------------------------------------
def _make_dm_env():
  m = types.ModuleType('dm_env')
  specs = types.ModuleType('dm_env.specs')

  class Array:
    def __init__(self, shape, dtype, name=None):
      self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

    def __repr__(self):
      return f'Array(shape={self.shape}, dtype={self.dtype}, name={self.name!r})'

  class BoundedArray(Array):
    def __init__(self, shape, dtype, minimum, maximum, name=None):
      super().__init__(shape, dtype, name)
      self.minimum, self.maximum = np.asarray(minimum, dtype=dtype), np.asarray(maximum, dtype=dtype)

  specs.Array, specs.BoundedArray = Array, BoundedArray

  class StepType(enum.IntEnum):
    FIRST = 0
    MID = 1
    LAST = 2

  class TimeStep(collections.namedtuple('TimeStep', ['step_type', 'reward', 'discount', 'observation'])):
    __slots__ = ()

    def first(self):
      return self.step_type == StepType.FIRST

    def mid(self):
      return self.step_type == StepType.MID

    def last(self):
      return self.step_type == StepType.LAST

  class Environment:
    pass

  m.specs, m.StepType, m.TimeStep, m.Environment = specs, StepType, TimeStep, Environment
  m.restart = lambda observation: TimeStep(StepType.FIRST, None, None, observation)
  m.transition = lambda reward, observation, discount=1.0: TimeStep(StepType.MID, reward, discount, observation)
  m.termination = lambda reward, observation: TimeStep(StepType.LAST, reward, 0.0, observation)
  m.truncation = lambda reward, observation, discount=1.0: TimeStep(StepType.LAST, reward, discount, observation)
  return m, specs


# ---------------------------------------------------------------------------------------------------------------
# lxml.etree stand-in (absent from this image): the handful of calls suite/cartpole.py, suite/quadruped.py and
# utils/xml_tools.py make — fromstring / XML / XMLParser / Element / SubElement / tostring, find / findall with
# [@name='..'] predicates, getparent() — on top of xml.etree.ElementTree
# ---------------------------------------------------------------------------------------------------------------
def _make_lxml():
  import xml.etree.ElementTree as ET

  class Element(ET.Element):
    _parent = None

    def getparent(self):
      return self._parent

    def append(self, child):
      super().append(child); child._parent = self

    def insert(self, index, child):
      super().insert(index, child); child._parent = self

    def remove(self, child):
      super().remove(child); child._parent = None

  def _link(root):
    for parent in root.iter():
      for child in parent:
        child._parent = parent
    return root

  class XMLParser:
    def __init__(self, remove_blank_text=False, **unused):
      self.remove_blank_text = remove_blank_text

  def fromstring(text, parser=None):
    builder = ET.TreeBuilder(element_factory=Element)
    p = ET.XMLParser(target=builder)
    p.feed(text)
    root = _link(p.close())
    if parser is not None and parser.remove_blank_text:
      for e in root.iter():
        if e.text is not None and not e.text.strip(): e.text = None
        if e.tail is not None and not e.tail.strip(): e.tail = None
    return root

  def SubElement(parent, tag, attrib=None, **extra):
    child = Element(tag, dict(attrib or {}, **extra))
    parent.append(child)
    return child

  def tostring(element, pretty_print=False, **unused):
    if pretty_print:
      ET.indent(element)
    return ET.tostring(element)

  lxml = types.ModuleType('lxml')
  etree = types.ModuleType('lxml.etree')
  etree.Element, etree.SubElement, etree.XMLParser = Element, SubElement, XMLParser
  etree.fromstring = etree.XML = fromstring
  etree.tostring = tostring
  class _Tree:
    def __init__(self, root): self._root = root
    def getroot(self): return self._root
    def find(self, path): return self._root.find(path)
    def findall(self, path): return self._root.findall(path)
    def iter(self, *a): return self._root.iter(*a)
  etree.parse = lambda file_obj, parser=None: _Tree(fromstring(file_obj.read(), parser))
  etree.tostring = lambda element, pretty_print=False, **kw: tostring(element.getroot() if isinstance(element, _Tree) else element, pretty_print)
  lxml.etree = etree
  return lxml, etree


def _package(name, path):
  p = types.ModuleType(name)
  p.__path__ = [path]
  p.__package__ = name
  return p


def install_suite_package():
  """After install(): make `dm_control.suite` the REAL package — execute the reference's own suite/__init__.py (which imports
  all 19 domain files and builds ALL_TASKS / BENCHMARKING / `suite.load`) — and provide what the reference's own
  suite_test.py imports besides (`mock`, `mjbindings.constants`, the mjtConstraint names suite/dog.py reads at import)."""
  import unittest.mock
  sys.modules.setdefault('mock', unittest.mock)
  mjb = sys.modules['dm_control.mujoco.wrapper.mjbindings']
  mjb.enums.mjtConstraint = types.SimpleNamespace(mjCNSTR_EQUALITY=0, mjCNSTR_FRICTION_DOF=1, mjCNSTR_FRICTION_TENDON=2,
                                                  mjCNSTR_LIMIT_JOINT=3, mjCNSTR_LIMIT_TENDON=4, mjCNSTR_CONTACT_FRICTIONLESS=5,
                                                  mjCNSTR_CONTACT_PYRAMIDAL=6, mjCNSTR_CONTACT_ELLIPTIC=7)
  consts = types.ModuleType('dm_control.mujoco.wrapper.mjbindings.constants')
  consts.mjMAXVAL, consts.mjMINVAL = 1e10, 1e-15
  mjb.constants = consts
  sys.modules['dm_control.mujoco.wrapper.mjbindings.constants'] = consts
  pkg = sys.modules['dm_control.suite']
  if not hasattr(pkg, 'ALL_TASKS'):
    init = os.path.join(REFERENCE, 'dm_control', 'suite', '__init__.py')
    pkg.__file__ = init
    exec(compile(open(init).read(), init, 'exec'), pkg.__dict__)
  return pkg


def install():
  if not available():
    raise RuntimeError(f'{REFERENCE}/dm_control not found')
  if 'dm_control.mujoco' in sys.modules and getattr(sys.modules['dm_control.mujoco'], '_b200_shim', False):
    return
  root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  if root not in sys.path:
    sys.path.insert(0, root)
  dm_env, specs = _make_dm_env()
  sys.modules['dm_env'], sys.modules['dm_env.specs'] = dm_env, specs
  if 'lxml' not in sys.modules:
    try:
      import lxml.etree      # noqa: F401  (the real one, where installed)
    except ImportError:
      sys.modules['lxml'], sys.modules['lxml.etree'] = _make_lxml()
  ref = os.path.join(REFERENCE, 'dm_control')
  for name, sub in (('dm_control', ''), ('dm_control.rl', 'rl'), ('dm_control.suite', 'suite'), ('dm_control.suite.utils', 'suite/utils'),
                    ('dm_control.utils', 'utils')):
    sys.modules[name] = _package(name, os.path.join(ref, sub))
  from dm_control_b200 import refview
  from dm_control_b200.physics import PhysicsError

  mj = types.ModuleType('dm_control.mujoco')
  mj._b200_shim = True
  mj.Physics = refview.SingleEnvPhysics

  def action_spec(physics):       # dm_control/mujoco/engine.py:948-966
    m = physics.model
    nu = m.nu
    is_limited = np.asarray(m.actuator_ctrllimited).ravel().astype(bool)
    rng = np.asarray(m.actuator_ctrlrange).reshape(nu, 2)
    minima = np.full(nu, -np.inf); maxima = np.full(nu, np.inf)
    minima[is_limited], maxima[is_limited] = rng[is_limited, 0], rng[is_limited, 1]
    return specs.BoundedArray(shape=(nu,), dtype=float, minimum=minima, maximum=maxima)
  mj.action_spec = action_spec
  wrapper = types.ModuleType('dm_control.mujoco.wrapper')
  mjb = types.ModuleType('dm_control.mujoco.wrapper.mjbindings')
  enums = types.SimpleNamespace(mjtJoint=types.SimpleNamespace(mjJNT_FREE=0, mjJNT_BALL=1, mjJNT_SLIDE=2, mjJNT_HINGE=3),
                                mjtSensor=types.SimpleNamespace(mjSENS_TOUCH=0, mjSENS_ACCELEROMETER=1, mjSENS_VELOCIMETER=2, mjSENS_GYRO=3,
                                                                mjSENS_FORCE=4, mjSENS_TORQUE=5, mjSENS_MAGNETOMETER=6, mjSENS_RANGEFINDER=7))
  mjb.enums = enums
  # suite/quadruped.py binds the name at import (only its `escape` task, hfield upload, calls into it); the randomizers
  # and their test use two quaternion helpers of MuJoCo's C API, restated here (mju_axisAngle2Quat, mju_rotVecQuat)
  def mju_axisAngle2Quat(res, axis, angle):
    s = np.sin(0.5 * angle)
    res[0] = np.cos(0.5 * angle); res[1:4] = np.asarray(axis, dtype=float) * s

  def mju_rotVecQuat(res, vec, quat):
    w, x, y, z = quat
    R = np.array([[w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
                  [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
                  [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])
    res[:] = R @ np.asarray(vec, dtype=float)
  mjb.mjlib = types.SimpleNamespace(mju_axisAngle2Quat=mju_axisAngle2Quat, mju_rotVecQuat=mju_rotVecQuat)
  wrapper.mjbindings = mjb
  mj.wrapper = wrapper
  sys.modules['dm_control.mujoco'] = mj
  sys.modules['dm_control.mujoco.wrapper'] = wrapper
  sys.modules['dm_control.mujoco.wrapper.mjbindings'] = mjb
  sys.modules['dm_control'].mujoco = mj
  # rl/control.py defines its own PhysicsError; the engine raises dm_control_b200's: make them one class once imported
  import dm_control.rl.control as control
  control.PhysicsError = PhysicsError
