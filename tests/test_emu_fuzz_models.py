"""Random models from the supported MJCF subset: the CUDA kernel source (under the CPU emulation, tests/emu) against
the oracle. The benchmark fixtures exercise hinge/free joints, capsules and spheres; this sweeps the rest of what the
compiler accepts — ball and slide joints, every primitive geom, springs, armature, limits, all four actuator
shortcuts with force/ctrl limits and filter dynamics, fixed tendons, joint and tendon equalities, every sensor kind,
both integrators, odd time steps and disable flags — on trees of random shape, with contacts.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import b200mj_emu as emu   # noqa: E402

from dm_control_b200 import mjcf_compile   # noqa: E402

pytestmark = pytest.mark.timeout(900)


def _f(v):
  return ' '.join(f'{x:.6g}' for x in np.atleast_1d(v))


class Gen:
  """One random model. Everything is drawn from `rs`, so a failing seed reproduces."""

  def __init__(self, seed):
    self.rs = np.random.RandomState(seed)
    self.joints, self.sites, self.bodies, self.nbody = [], [], [], 0

  def geom(self, allow_all=True):
    rs = self.rs
    kind = rs.choice(['sphere', 'capsule', 'box', 'ellipsoid', 'cylinder'] if allow_all else ['sphere', 'capsule'],
                     p=[.3, .35, .15, .1, .1] if allow_all else [.5, .5])
    attrs = f'type="{kind}" '
    if kind == 'sphere':
      attrs += f'size="{rs.uniform(.04, .1):.4g}" pos="{_f(rs.uniform(-.05, .05, 3))}"'
    elif kind == 'capsule':
      if rs.rand() < .5:
        attrs += f'size="{rs.uniform(.03, .06):.4g}" fromto="0 0 0 {_f(rs.uniform(-.25, .25, 3))}"'
      else:
        attrs += f'size="{rs.uniform(.03, .06):.4g} {rs.uniform(.05, .15):.4g}" euler="{_f(rs.uniform(-90, 90, 3))}"'
    elif kind == 'box':
      attrs += f'size="{_f(rs.uniform(.04, .1, 3))}" euler="{_f(rs.uniform(-30, 30, 3))}"'
    elif kind == 'ellipsoid':
      attrs += f'size="{_f(rs.uniform(.04, .1, 3))}"'
    else:
      attrs += f'size="{rs.uniform(.04, .08):.4g} {rs.uniform(.04, .1):.4g}" euler="{_f(rs.uniform(-60, 60, 3))}"'
    if rs.rand() < .3:
      attrs += f' condim="{rs.choice([1, 3])}"'
    if rs.rand() < .3:
      attrs += f' friction="{rs.uniform(.3, 1.2):.3g} .005 .0001"'
    if rs.rand() < .2:
      attrs += f' density="{rs.uniform(300, 3000):.4g}"'
    if rs.rand() < .2:
      attrs += f' solref="{rs.uniform(.01, .04):.3g} {rs.uniform(.7, 1.3):.3g}"'
    if rs.rand() < .15:
      attrs += f' margin="{rs.uniform(0, .02):.3g}"'
    return f'<geom {attrs}/>'

  def joint(self, name, allow_ball=True):
    rs = self.rs
    kind = rs.choice(['hinge', 'slide', 'ball'], p=[.6, .2, .2])
    if kind == 'ball' and not allow_ball:
      kind = 'hinge'
    a = f'name="{name}" type="{kind}"'
    if kind != 'ball':
      axis = rs.randn(3); axis /= np.linalg.norm(axis)
      a += f' axis="{_f(axis)}"'
      if rs.rand() < .5:
        lo, hi = (-rs.uniform(20, 90), rs.uniform(20, 90)) if kind == 'hinge' else (-rs.uniform(.05, .2), rs.uniform(.05, .2))
        a += f' limited="true" range="{lo:.4g} {hi:.4g}"'
      if rs.rand() < .4:
        a += f' stiffness="{rs.uniform(1, 30):.4g}" springref="{rs.uniform(-10, 10) if kind == "hinge" else rs.uniform(-.05, .05):.4g}"'
      self.joints.append(name)
    if rs.rand() < .7:
      a += f' damping="{rs.uniform(.05, 2):.4g}"'
    if rs.rand() < .3:
      a += f' armature="{rs.uniform(.001, .05):.4g}"'
    if rs.rand() < .3:
      a += f' pos="{_f(rs.uniform(-.05, .05, 3))}"'
    return f'<joint {a}/>'

  def body(self, depth, free_root):
    rs = self.rs
    name = f'b{self.nbody}'; self.nbody += 1
    self.bodies.append(name)
    pos = rs.uniform(-.25, .25, 3) if depth else np.array([rs.uniform(-1, 1), rs.uniform(-1, 1), rs.uniform(.25, .8)])
    out = f'<body name="{name}" pos="{_f(pos)}"'
    if rs.rand() < .4:
      out += f' euler="{_f(rs.uniform(-40, 40, 3))}"'
    out += '>'
    if depth == 0 and free_root:
      out += f'<freejoint name="{name}_free"/>'
    else:
      # a ball joint stays alone on its body: together with another rotational joint at the same anchor the
      # joint-space inertia would be singular
      first = self.joint(f'{name}_j0')
      out += first
      if 'type="ball"' not in first and rs.rand() < .33:
        out += self.joint(f'{name}_j1', allow_ball=False)
    for _ in range(rs.choice([1, 1, 2])):
      out += self.geom()
    if rs.rand() < .6:
      sname = f's{len(self.sites)}'; self.sites.append(sname)
      out += f'<site name="{sname}" pos="{_f(rs.uniform(-.05, .05, 3))}" size="{rs.uniform(.05, .15):.3g}" type="{rs.choice(["sphere", "box", "capsule"])}"/>'
    if depth < 3:
      for _ in range(rs.choice([0, 1, 1, 2]) if depth else rs.choice([1, 2])):
        if self.nbody < 9:
          out += self.body(depth + 1, free_root)
    return out + '</body>'

  def xml(self):
    rs = self.rs
    trees = ''.join(self.body(0, rs.rand() < .6) for _ in range(rs.choice([1, 2])))
    opt = f'<option timestep="{rs.choice([.002, .004, .005, .01])}" integrator="{rs.choice(["Euler", "Euler", "RK4"])}"'
    if rs.rand() < .2:
      opt += f' gravity="{_f(rs.uniform(-3, 3, 2))} -9.81"'
    opt += '>'
    flags = [f for f in ('eulerdamp', 'filterparent', 'warmstart', 'refsafe', 'clampctrl') if rs.rand() < .12]
    if flags:
      opt += '<flag ' + ' '.join(f'{f}="disable"' for f in flags) + '/>'
    opt += '</option>'
    acts, tend, eqs, sens = [], [], [], []
    sj = list(self.joints)
    rs.shuffle(sj)
    for k, j in enumerate(sj[:6]):
      kind = rs.choice(['motor', 'position', 'velocity', 'general'])
      a = f'name="a{k}" joint="{j}"'
      if kind == 'motor':
        a += f' gear="{rs.uniform(1, 20):.4g}"'
      elif kind == 'position':
        a += f' kp="{rs.uniform(2, 40):.4g}"'
      elif kind == 'velocity':
        a += f' kv="{rs.uniform(.2, 3):.4g}"'
      else:
        dyn = rs.choice(['none', 'filter', 'integrator'], p=[.3, .5, .2])
        a += (f' gainprm="{rs.uniform(1, 10):.4g}" biastype="affine" biasprm="{_f(rs.uniform(-1, 1, 3))}"'
              + (f' dyntype="filter" dynprm="{rs.uniform(.02, .2):.3g}"' if dyn == 'filter' else '')
              + (' dyntype="integrator" actlimited="true" actrange="-1 1"' if dyn == 'integrator' else ''))
      if rs.rand() < .6:
        a += ' ctrllimited="true" ctrlrange="-.8 .9"'
      if rs.rand() < .3:
        a += f' forcelimited="true" forcerange="-{rs.uniform(1, 10):.3g} {rs.uniform(1, 10):.3g}"'
      acts.append(f'<{kind} {a}/>')
    if len(sj) >= 2 and rs.rand() < .6:
      for k in range(rs.choice([1, 2])):
        js = rs.choice(sj, size=min(len(sj), rs.choice([2, 3])), replace=False)
        tend.append(f'<fixed name="t{k}">' + ''.join(f'<joint joint="{j}" coef="{rs.uniform(-1.5, 1.5):.3g}"/>' for j in js) + '</fixed>')
      if rs.rand() < .5:
        eqs.append(f'<tendon tendon1="t0" polycoef="{rs.uniform(-.05, .05):.3g} 1 0 0 0"/>' if len(tend) == 1 or rs.rand() < .5
                   else '<tendon tendon1="t0" tendon2="t1" polycoef="0 1 0 0 0"/>')
      if rs.rand() < .4:
        acts.append(f'<motor name="at" tendon="t0" gear="{rs.uniform(1, 5):.3g}"/>')
    if len(sj) >= 2 and rs.rand() < .4:
      eqs.append(f'<joint joint1="{sj[0]}" joint2="{sj[1]}" polycoef="0 {rs.uniform(.5, 1.5):.3g} 0 0 0"/>')
    for s in self.sites:
      for tag in ('touch', 'accelerometer', 'velocimeter', 'gyro', 'force', 'torque'):
        if rs.rand() < .35:
          sens.append(f'<{tag} site="{s}"/>')
    for j in sj[:3]:
      sens.append(f'<jointpos joint="{j}"/><jointvel joint="{j}"/>')
    for k in range(len(acts)):
      if rs.rand() < .3:
        sens.append(f'<actuatorfrc actuator="{acts[k].split(chr(34))[1]}"/>')
    for b in self.bodies[:2]:
      sens.append(f'<subtreecom body="{b}"/><subtreelinvel body="{b}"/>')
    if self.sites and len(self.bodies) > 1:
      sens.append(f'<framepos objtype="site" objname="{self.sites[0]}"/>'
                  f'<framepos objtype="xbody" objname="{self.bodies[-1]}" reftype="xbody" refname="{self.bodies[0]}"/>')
    return (f'<mujoco>{opt}<compiler angle="degree"/><worldbody><geom name="floor" type="plane" size="5 5 .1"'
            f' friction="{rs.uniform(.5, 1.2):.3g} .005 .0001"/>{trees}</worldbody>'
            f'<actuator>{"".join(acts)}</actuator><tendon>{"".join(tend)}</tendon><equality>{"".join(eqs)}</equality>'
            f'<sensor>{"".join(sens)}</sensor></mujoco>')


def relerr(a, b):
  a, b = np.asarray(a), np.asarray(b)
  return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))) if a.size else 0.0


@pytest.mark.parametrize('seed', range(120))
def test_random_model_rollout(seed, oracle_mod):
  xml = Gen(seed).xml()
  model = mjcf_compile.compile_xml(xml)
  if model.nv == 0:
    pytest.skip('degenerate draw')
  rs = np.random.RandomState(1000 + seed)
  B, nsteps = 3, 40
  p = emu.EmuPhysics(model, B)
  oracles = [oracle_mod.OraclePhysics(model) for _ in range(B)]
  v0 = rs.uniform(-1, 1, (B, model.nv))
  p.data.qvel[:] = v0
  p.forward()
  for e, o in enumerate(oracles):
    o.qvel[:] = v0[e]; o.forward()
  tape = rs.uniform(-1.2, 1.2, (nsteps, B, model.nu))
  act0 = None
  ncon_seen = 0
  for t in range(nsteps):
    n = int(rs.choice([1, 1, 2, 3]))
    p.data.ctrl[:] = tape[t]
    p.step(n)
    for e, o in enumerate(oracles):
      o.ctrl[:] = tape[t, e]
      o.control_step(n)
      if o.warning.any():           # diverged draw (e.g. conflicting equalities): both sides must agree that it did
        assert p.data.warning[e].any(), (seed, t, e)
        return
      if np.abs(o.qvel).max() > 100:      # a violently unstable draw amplifies rounding differences without bound: stop here
        return
      # Contacts of pairs without a closed form (MPR: an ellipsoid, a non-plane cylinder or two boxes involved) are a
      # discontinuous function of the pose — portal and start-direction choices flip under 1e-16 perturbations — and the
      # random trees start with such geoms deeply interpenetrating: once one is active, rounding differences between
      # kernel and oracle no longer stay small. Those pairs have their own tests (tests/test_convex_pairs.py, the
      # convex_zoo golden); here the comparison ends at the first step that has one.
      def _mpr(c):
        t1, t2 = int(model.geom_type[c.geom1]), int(model.geom_type[c.geom2])
        return t1 != 0 and (t1 in (4, 5) or t2 in (4, 5) or (t1 == 6 and t2 == 6))
      if any(_mpr(c) for c in o.contact):
        return
      # the north-star bar (1e-5); agreement starts near 1e-13 and random mechanisms thrashing under random controls
      # amplify it chaotically over the 40 calls
      assert relerr(p.data.qpos[e], o.qpos) < 1e-6 and relerr(p.data.qvel[e], o.qvel) < 1e-5, (seed, t, e, xml)
      assert int(p.data.ncon[e]) == o.ncon and int(p.data.nefc[e]) == o.nefc, (seed, t, e)
      assert [tuple(x) for x in p.data.contact_geom[e, :o.ncon]] == [(c.geom1, c.geom2) for c in o.contact], (seed, t, e)
      if model.na:
        assert relerr(p.data.act[e], o.act) < 1e-7
      if model.nsensordata:
        o.subtree_vel()
        assert relerr(p.data.sensordata[e], o.sensordata) < 1e-5, (seed, t, e)
      ncon_seen += o.ncon
  assert not p.data.warning.any()


_ORDER_SCRIPT = r'''
import hashlib, os, sys
import numpy as np
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'tests')); sys.path.insert(0, os.path.join(root, 'tests', 'emu'))
import b200mj_emu as emu
from test_emu_fuzz_models import Gen
from dm_control_b200 import mjcf_compile
h = hashlib.sha256()
for seed in range(0, 120, 6):
  model = mjcf_compile.compile_xml(Gen(seed).xml())
  rs = np.random.RandomState(1000 + seed)
  p = emu.EmuPhysics(model, 2)
  p.data.qvel[:] = rs.uniform(-1, 1, (2, model.nv)); p.forward()
  for t in range(12):
    p.data.ctrl[:] = rs.uniform(-1, 1, (2, model.nu)); p.step(int(rs.choice([1, 2])))
    for f in ('qpos', 'qvel', 'sensordata', 'ncon', 'contact_geom', 'qacc', 'efc_force', 'act'):
      h.update(np.ascontiguousarray(getattr(p.data, f)).tobytes())
print(h.hexdigest())
'''


def test_lane_order_does_not_change_results():
  """Race check without a GPU: 20 random models stepped with the emulator scheduling the lanes of each block in
  ascending and in descending order must agree bit for bit (see cuda_emu.h: emu_dir)."""
  import subprocess
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  def digest(**env):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, '-c', _ORDER_SCRIPT, root], env=e, capture_output=True, text=True, timeout=800)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()[-1]
  assert digest() == digest(B200MJ_EMU_ORDER='reverse')
