"""Model fixtures + seeded initial states shared by tests, bench.py and smoke().

`load(name)` returns a compiled Model: either one of the hot-path configs (compiled fixtures under assets/, see
tools/make_model_fixtures.py) or one of the small hand-written MJCF strings below (this repo's own test models,
exercising joint kinds / geom pairs the suite models do not).
"""
from __future__ import annotations

import os

import numpy as np

from . import mjcf_compile
from .model import Model

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')

XML = {
    # README known-answer model (dm_control/mujoco/README.md:11-24 geometry): slide joint, box + sphere
    'slide_box': """
<mujoco><worldbody>
  <geom name="floor" type="plane" size="1 1 .1"/>
  <body name="box" pos="0 0 .3">
    <joint name="up_down" type="slide" axis="0 0 1"/>
    <geom name="box" type="box" size=".2 .2 .2"/>
    <geom name="sphere" pos=".2 .2 .2" size=".1"/>
  </body>
</worldbody></mujoco>""",
    # free box on a plane (wrapper/core_test.py:393-416 weight test geometry)
    'free_box': """
<mujoco><worldbody>
  <geom name="floor" type="plane" size="1 1 .1"/>
  <body name="box" pos="0 0 .1"><freejoint/><geom name="box" type="box" size=".1 .1 .1"/></body>
</worldbody></mujoco>""",
    # two free capsules + a free sphere falling on a plane, colliding with each other; a hinged pendulum chain
    'pendulum_free': """
<mujoco><option timestep="0.004"/>
<default><geom friction=".8" solref=".02 1"/></default>
<worldbody>
  <geom name="floor" type="plane" size="5 5 .1"/>
  <body name="a" pos="0 0 .5"><freejoint name="a"/><geom name="a" type="capsule" size=".08 .2" euler="0 80 0"/></body>
  <body name="b" pos=".05 .02 .9"><freejoint name="b"/><geom name="b" type="capsule" size=".06 .25" euler="70 0 20"/></body>
  <body name="s" pos="-.1 .05 1.3"><freejoint name="s"/><geom name="s" type="sphere" size=".12" condim="1"/></body>
  <body name="p1" pos="1 0 1"><joint name="h1" type="hinge" axis="0 1 0" damping=".1" range="-60 60" limited="true"/>
    <geom name="p1" type="capsule" fromto="0 0 0 0 0 -.4" size=".04"/>
    <body name="p2" pos="0 0 -.4"><joint name="h2" type="hinge" axis="1 0 0" stiffness="2" springref="20"/>
      <geom name="p2" type="capsule" fromto="0 0 0 .3 0 -.3" size=".03"/>
      <body name="p3" pos=".3 0 -.3"><joint name="sl" type="slide" axis="0 0 1" range="-.1 .1" limited="true" damping="1"/>
        <geom name="p3" type="sphere" size=".06"/></body></body></body>
</worldbody>
<actuator><motor name="m1" joint="h1" gear="3" ctrllimited="true" ctrlrange="-1 1"/>
  <position name="m2" joint="h2" kp="5"/></actuator>
<sensor><jointpos joint="h1"/><jointvel joint="h2"/><subtreecom body="p1"/><subtreelinvel body="p1"/></sensor>
</mujoco>""",
}

# every convex primitive against every other and against a static box platform: ellipsoid / cylinder / box / capsule /
# sphere free bodies dropped in a heap (pairs resolved by include/b200mj_convex.h: MPR and capsule-box)
XML['convex_zoo'] = """
<mujoco><option timestep="0.004"/>
<default><geom friction=".7" solref=".02 1"/></default>
<worldbody>
  <geom name="floor" type="plane" size="5 5 .1"/>
  <geom name="platform" type="box" pos="0 0 .15" size=".6 .5 .15"/>
  <body name="e" pos="0 0 .55"><freejoint/><geom name="e" type="ellipsoid" size=".18 .12 .08" euler="20 30 0"/></body>
  <body name="c" pos=".05 .05 .85"><freejoint/><geom name="c" type="cylinder" size=".09 .12" euler="70 10 0"/></body>
  <body name="b" pos="-.05 .02 1.15"><freejoint/><geom name="b" type="box" size=".1 .08 .06" euler="10 40 25"/></body>
  <body name="k" pos=".02 -.04 1.45"><freejoint/><geom name="k" type="capsule" size=".05 .14" euler="80 0 30"/></body>
  <body name="s" pos="0 .03 1.7"><freejoint/><geom name="s" type="sphere" size=".09"/></body>
  <body name="k2" pos=".45 .3 .6"><freejoint/><geom name="k2" type="capsule" size=".04 .2" euler="90 0 60"/></body>
</worldbody>
</mujoco>"""

_CACHE = {}


def load(name):
  if name.endswith('_floor'):        # same model, other family of start states (initial_states)
    name = name[:-len('_floor')]
  if name in _CACHE:
    return _CACHE[name]
  if name in XML:
    m = mjcf_compile.compile_xml(XML[name])
  else:
    path = os.path.join(_ASSETS, name + '.npz')
    if not os.path.exists(path):
      raise FileNotFoundError(f'{path}: run tools/make_model_fixtures.py (needs /root/reference)')
    m = Model.load(path)
  _CACHE[name] = m
  return m


def initial_states(model, name, batch, seed=0):
  """Seeded [B, nq], [B, nv] start states: env i uses RandomState(seed*100003 + i).

  Limited hinge/slide joints uniform in range, free-joint orientation from `rand(4)` normalised (the reference's
  randomizer convention, suite/utils/randomizers.py:35-88), root height spread so some envs start in contact.
  """
  nq, nv = model.nq, model.nv
  q = np.tile(model.qpos0, (batch, 1)).astype(np.float64)
  v = np.zeros((batch, nv))
  if name.endswith('_floor'):
    # upright, a little above the floor, small joint offsets, at rest: the model drops onto its feet / toes and stays
    # in contact (the tumbling starts below launch a deeply penetrating quadruped into the air instead)
    for e in range(batch):
      rs = np.random.RandomState(seed * 100003 + e + 555)
      for j in range(model.njnt):
        t, qa = model.jnt_type[j], model.jnt_qposadr[j]
        if t == 0:
          q[e, qa + 2] = model.qpos0[qa + 2] + rs.uniform(0.02, 0.12)
        elif t in (2, 3) and model.jnt_limited[j]:
          lo, hi = model.jnt_range[j]
          q[e, qa] = np.clip(model.qpos0[qa] + rs.uniform(-0.1, 0.1), lo + 0.02 * (hi - lo), hi - 0.02 * (hi - lo))
    return q, v
  for e in range(batch):
    rs = np.random.RandomState(seed * 100003 + e)
    for j in range(model.njnt):
      t, qa, da = model.jnt_type[j], model.jnt_qposadr[j], model.jnt_dofadr[j]
      lo, hi = model.jnt_range[j]
      if t in (2, 3):
        if model.jnt_limited[j]:
          # stay inside the middle 80% of the range so the start is not already at a limit
          q[e, qa] = rs.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo))
        elif t == 3:
          q[e, qa] = rs.uniform(-np.pi, np.pi)
        v[e, da] = rs.uniform(-0.5, 0.5)
      elif t == 0:
        quat = rs.rand(4)
        q[e, qa + 3:qa + 7] = quat / np.linalg.norm(quat)
        q[e, qa + 2] = q[e, qa + 2] * rs.uniform(0.35, 1.0)
        v[e, da:da + 6] = rs.uniform(-0.5, 0.5, 6)
  if name == 'cartpole':
    q[:, 0] = np.clip(q[:, 0], -1.0, 1.0)
  if name == 'cmu_humanoid':
    # 56 joints drawn over their whole range fold the walker into itself (hundreds of constraint rows): start
    # from the upright pose with small joint offsets and a slightly lowered root instead
    for e in range(batch):
      rs = np.random.RandomState(seed * 100003 + e + 77)
      q[e] = model.qpos0
      q[e, 2] = model.qpos0[2] * rs.uniform(0.93, 1.0)
      for j in range(model.njnt):
        if model.jnt_type[j] == 3:
          lo, hi = model.jnt_range[j]
          qa = model.jnt_qposadr[j]
          q[e, qa] = np.clip(model.qpos0[qa] + rs.uniform(-0.15, 0.15), lo + 0.02, hi - 0.02)
  return q, v
