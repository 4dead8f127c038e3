"""MJCF-subset compiler (`dm_control_b200/mjcf_compile.py`) against MuJoCo's documented compile semantics (CPU).

The compiler stands where `mujoco.MjModel.from_xml_string` stands (`dm_control/mujoco/wrapper/core.py:179-182`); the
tables it emits feed the CUDA kernel and the oracle alike, so an error here would be invisible to the parity tests.
Checked here: closed-form mass/inertia of every primitive, composite inertial frames, `fromto`, default classes,
angle units, `settotalmass`, actuator shortcuts, option defaults, collision-pair filtering, tree tables, and the model
sizes SURVEY.md §8 counted from the reference XMLs.
"""
import math

import numpy as np
import pytest

from dm_control_b200 import mjcf_compile, testing_models as tm

RHO = 1000.0


def compile_(body_xml, extra='', top=''):
  return mjcf_compile.compile_xml(f'<mujoco>{top}<worldbody>{body_xml}</worldbody>{extra}</mujoco>')


@pytest.mark.parametrize('geom,mass,inertia', [
    ('type="sphere" size=".1"', RHO * 4 / 3 * math.pi * .1 ** 3, None),
    ('type="box" size=".1 .2 .3"', RHO * 8 * .1 * .2 * .3, 'box'),
    ('type="cylinder" size=".1 .25"', RHO * math.pi * .1 ** 2 * .5, 'cyl'),
    ('type="capsule" size=".1 .25"', RHO * (math.pi * .1 ** 2 * .5 + 4 / 3 * math.pi * .1 ** 3), 'cap'),
    ('type="ellipsoid" size=".1 .2 .3"', RHO * 4 / 3 * math.pi * .1 * .2 * .3, 'ell'),
])
def test_primitive_mass_and_inertia(geom, mass, inertia):
  m = compile_(f'<body name="b"><freejoint/><geom {geom}/></body>')
  assert m.body_mass[1] == pytest.approx(mass, rel=1e-12)
  got = np.sort(np.asarray(m.body_inertia[1]))
  if inertia is None:
    want = np.full(3, 0.4 * mass * .01)
  elif inertia == 'box':
    a, b, c = .1, .2, .3
    want = mass / 3 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
  elif inertia == 'cyl':
    r, h = .1, .25
    want = np.array([mass * (r * r / 4 + h * h / 3)] * 2 + [mass * r * r / 2])
  elif inertia == 'cap':
    r, h = .1, .25
    mc, ms = RHO * math.pi * r * r * 2 * h, RHO * 4 / 3 * math.pi * r ** 3
    ixx = mc * (r * r / 4 + h * h / 3) + ms * (0.4 * r * r + h * h + 0.75 * h * r)
    want = np.array([ixx, ixx, mc * r * r / 2 + ms * 0.4 * r * r])
  else:
    a, b, c = .1, .2, .3
    want = mass / 5 * np.array([b * b + c * c, a * a + c * c, a * a + b * b])
  np.testing.assert_allclose(got, np.sort(want), rtol=1e-12)


def test_composite_inertial_frame_parallel_axis():
  d, r = .3, .1
  m = compile_(f'<body name="b" pos="1 2 3"><freejoint/><geom type="sphere" size="{r}" pos="{d} 0 0"/>'
               f'<geom type="sphere" size="{r}" pos="-{d} 0 0" density="3000"/></body>')
  m1, m2 = RHO * 4 / 3 * math.pi * r ** 3, 3 * RHO * 4 / 3 * math.pi * r ** 3
  assert m.body_mass[1] == pytest.approx(m1 + m2)
  com = (m1 * d - m2 * d) / (m1 + m2)
  np.testing.assert_allclose(m.body_ipos[1], [com, 0, 0], atol=1e-14)
  i_sph = 0.4 * (m1 + m2) * r * r
  i_off = i_sph + m1 * (d - com) ** 2 + m2 * (d + com) ** 2
  np.testing.assert_allclose(np.sort(m.body_inertia[1]), np.sort([i_sph, i_off, i_off]), rtol=1e-12)
  assert m.body_subtreemass[0] == pytest.approx(m1 + m2) and m.body_subtreemass[1] == pytest.approx(m1 + m2)


def test_fromto_capsule_and_explicit_inertial():
  m = compile_('<body name="b"><joint type="hinge" axis="0 1 0"/>'
               '<geom name="g" type="capsule" fromto="0 0 0 .3 0 -.4" size=".05"/></body>'
               '<body name="c" pos="0 1 0"><inertial pos=".1 0 0" mass="2.5" diaginertia=".3 .2 .1"/>'
               '<joint type="slide" axis="1 0 0"/><geom type="sphere" size=".1"/></body>')
  g = m.name2id('g', 'geom')
  np.testing.assert_allclose(m.geom_pos[g], [.15, 0, -.2], atol=1e-15)
  assert m.geom_size[g][0] == pytest.approx(.05) and m.geom_size[g][1] == pytest.approx(.25)
  # the capsule's local z axis points along the segment (either sense: the capsule is symmetric)
  w, x, y, z = m.geom_quat[g]
  zaxis = np.array([2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y)])
  np.testing.assert_allclose(np.abs(zaxis @ np.array([.6, 0, -.8])), 1.0, atol=1e-12)
  c = m.name2id('c', 'body')
  assert m.body_mass[c] == 2.5                                       # <inertial> overrides the geoms
  np.testing.assert_allclose(m.body_ipos[c], [.1, 0, 0]); np.testing.assert_allclose(m.body_inertia[c], [.3, .2, .1])


def test_defaults_classes_angles_and_settotalmass():
  m = mjcf_compile.compile_xml("""
<mujoco>
  <compiler angle="degree" settotalmass="10"/>
  <default>
    <joint type="hinge" axis="0 1 0" damping=".5" limited="true" range="-90 45"/>
    <geom type="capsule" size=".05 .2" friction=".7 .1 .1"/>
    <default class="arm"><joint damping="2" range="-30 30"/><geom size=".03 .1"/></default>
  </default>
  <worldbody>
    <body name="a"><joint name="ja"/><geom name="ga"/>
      <body name="b" pos="0 0 -.5" childclass="arm"><joint name="jb"/><geom name="gb"/>
        <body name="c" pos="0 0 -.3"><joint name="jc" damping="7"/><geom name="gc" class="main" type="sphere" size=".04"/></body>
      </body>
    </body>
  </worldbody>
</mujoco>""")
  ja, jb, jc = (m.name2id(n, 'joint') for n in ('ja', 'jb', 'jc'))
  np.testing.assert_allclose(m.jnt_range[ja], np.radians([-90, 45]))
  np.testing.assert_allclose(m.jnt_range[jb], np.radians([-30, 30]))
  np.testing.assert_allclose(m.jnt_range[jc], np.radians([-30, 30]))           # childclass is inherited by grandchildren
  assert [m.dof_damping[m.jnt_dofadr[j]] for j in (ja, jb, jc)] == [.5, 2, 7]
  assert all(m.jnt_limited[j] for j in (ja, jb, jc))
  ga, gb, gc = (m.name2id(n, 'geom') for n in ('ga', 'gb', 'gc'))
  np.testing.assert_allclose(m.geom_size[ga][:2], [.05, .2]); np.testing.assert_allclose(m.geom_size[gb][:2], [.03, .1])
  assert m.geom_size[gc][0] == pytest.approx(.04)                              # explicit class="main" + own attributes
  np.testing.assert_allclose(m.geom_friction[gb], [.7, .1, .1])
  assert float(np.sum(m.body_mass)) == pytest.approx(10.0, rel=1e-12)          # settotalmass rescales masses...
  ratio = m.body_mass[1] / (RHO * (math.pi * .05 ** 2 * .4 + 4 / 3 * math.pi * .05 ** 3))
  r, h = .05, .2
  mc, ms = RHO * math.pi * r * r * 2 * h, RHO * 4 / 3 * math.pi * r ** 3
  izz = (mc * r * r / 2 + ms * 0.4 * r * r) * ratio
  assert np.min(m.body_inertia[1]) == pytest.approx(izz, rel=1e-12)            # ...and inertias by the same factor


def test_option_and_contact_parameter_defaults():
  m = compile_('<geom name="floor" type="plane" size="1 1 .1"/><body><freejoint/><geom name="s" type="sphere" size=".1"/></body>')
  assert m.opt.timestep == 0.002 and tuple(m.opt.gravity) == (0, 0, -9.81)
  assert m.opt.integrator == 0 and m.opt.tolerance == 1e-8 and m.opt.iterations == 100 and m.opt.ls_iterations == 50
  assert m.opt.impratio == 1 and m.opt.disableflags == 0
  s = m.name2id('s', 'geom')
  np.testing.assert_allclose(m.geom_friction[s], [1, .005, .0001])
  np.testing.assert_allclose(m.geom_solref[s], [.02, 1]); np.testing.assert_allclose(m.geom_solimp[s], [.9, .95, .001, .5, 2])
  assert m.geom_condim[s] == 3 and m.geom_contype[s] == 1 and m.geom_conaffinity[s] == 1 and m.geom_margin[s] == 0
  # free body: translational inverse weight 1/m, rotational 1/I (sphere: isotropic)
  mass = RHO * 4 / 3 * math.pi * .1 ** 3
  np.testing.assert_allclose(m.body_invweight0[1], [1 / mass, 1 / (0.4 * mass * .01)], rtol=1e-10)
  np.testing.assert_allclose(m.dof_invweight0, [1 / mass] * 3 + [1 / (0.4 * mass * .01)] * 3, rtol=1e-10)


def test_actuator_shortcuts():
  m = compile_('<body><joint name="j" type="hinge" axis="0 0 1"/><geom type="sphere" size=".1"/>'
               '<body pos="0 0 .3"><joint name="k" type="slide" axis="0 0 1"/><geom type="sphere" size=".1"/></body></body>',
               extra='<actuator><motor name="m" joint="j" gear="40" ctrllimited="true" ctrlrange="-1 1"/>'
                     '<position name="p" joint="k" kp="30"/><velocity name="v" joint="j" kv="3"/>'
                     '<general name="g" joint="k" dyntype="filter" dynprm=".1" gainprm="5" biastype="affine" biasprm="1 -2 -3"/>'
                     '</actuator>')
  mi, pi_, vi, gi = (m.name2id(n, 'actuator') for n in 'mpvg')
  assert m.actuator_gear[mi] == 40 and m.actuator_ctrllimited[mi] and tuple(m.actuator_ctrlrange[mi]) == (-1, 1)
  assert m.actuator_gainprm[mi][0] == 1 and not m.actuator_biasprm[mi].any()                   # motor: gain 1, no bias
  assert m.actuator_gainprm[pi_][0] == 30 and tuple(m.actuator_biasprm[pi_][:3]) == (0, -30, 0)  # position: kp, (0,-kp,0)
  assert m.actuator_gainprm[vi][0] == 3 and tuple(m.actuator_biasprm[vi][:3]) == (0, 0, -3)      # velocity: kv, (0,0,-kv)
  assert m.actuator_gainprm[gi][0] == 5 and tuple(m.actuator_biasprm[gi][:3]) == (1, -2, -3) and m.actuator_dynprm[gi] == .1
  assert m.na == 1 and m.actuator_actadr[gi] == 0 and m.actuator_actadr[mi] == -1
  assert m.actuator_trnid[pi_] == m.name2id('k', 'joint')


def test_collision_pair_filtering():
  xml = ('<geom name="floor" type="plane" size="1 1 .1"/>'
         '<body name="a" pos="0 0 1"><freejoint/><geom name="a1" type="sphere" size=".1"/><geom name="a2" type="sphere" size=".1" pos=".3 0 0"/>'
         '<body name="b" pos="0 0 .3"><joint type="hinge" axis="0 1 0"/><geom name="b1" type="capsule" size=".05 .1"/>'
         '<body name="c" pos="0 0 .3"><joint type="hinge" axis="0 1 0"/><geom name="c1" type="sphere" size=".05"/></body></body></body>'
         '<body name="d" pos="1 0 1"><freejoint/><geom name="d1" type="sphere" size=".1" contype="2" conaffinity="4"/></body>')
  m = compile_(xml)
  names = lambda i: m.id2name(int(i), 'geom')
  pairs = {frozenset((names(a), names(b))) for a, b in zip(m.pair_geom1, m.pair_geom2)}
  assert frozenset(('a1', 'a2')) not in pairs                       # same body
  assert frozenset(('a1', 'b1')) not in pairs and frozenset(('b1', 'c1')) not in pairs    # parent-child filter
  assert frozenset(('a1', 'c1')) in pairs                           # grandparent-grandchild do collide
  assert frozenset(('floor', 'a1')) in pairs and frozenset(('floor', 'c1')) in pairs
  assert not any('d1' in p for p in pairs)                          # contype 2 / conaffinity 4 matches nothing here
  m2 = compile_(xml, extra='<contact><exclude body1="a" body2="c"/></contact>')
  pairs2 = {frozenset((m2.id2name(int(a), 'geom'), m2.id2name(int(b), 'geom'))) for a, b in zip(m2.pair_geom1, m2.pair_geom2)}
  assert frozenset(('a1', 'c1')) not in pairs2 and frozenset(('floor', 'a1')) in pairs2
  m3 = compile_(xml, top='<option><flag filterparent="disable"/></option>')
  pairs3 = {frozenset((m3.id2name(int(a), 'geom'), m3.id2name(int(b), 'geom'))) for a, b in zip(m3.pair_geom1, m3.pair_geom2)}
  assert frozenset(('a1', 'b1')) in pairs3


def test_tree_tables():
  m = tm.load('humanoid')
  # parents precede children; dof_parentid walks towards the root inside a body then to the parent body's last dof
  assert all(m.body_parentid[b] < b for b in range(1, m.nbody))
  for i in range(m.nv):
    p = m.dof_parentid[i]
    assert p < i
    if p >= 0:
      bi, bp = m.dof_bodyid[i], m.dof_bodyid[p]
      assert bp == bi or bp == m.body_parentid[bi] or m.body_dofnum[m.body_parentid[bi]] == 0
  # levels partition the bodies by depth
  depth = np.zeros(m.nbody, int)
  for b in range(1, m.nbody):
    depth[b] = depth[m.body_parentid[b]] + 1
  for l in range(m.nlevel):
    bodies = m.level_body[m.level_adr[l]:m.level_adr[l + 1]]
    assert len(bodies) and all(depth[b] == l for b in bodies)
  assert m.level_adr[m.nlevel] == m.nbody
  # body_dofmask: bit i set iff dof i moves the body (dof's body is the body or one of its ancestors)
  for b in range(m.nbody):
    anc = set()
    x = b
    while x > 0:
      anc.add(x); x = m.body_parentid[x]
    for i in range(m.nv):
      bit = (int(np.asarray(m.body_dofmask).reshape(-1)[2 * b + (i >> 5)]) >> (i & 31)) & 1
      assert bit == (m.dof_bodyid[i] in anc)
  assert np.all(m.body_rootid[1:] == 1)


@pytest.mark.parametrize('name,sizes', [
    ('cartpole', dict(nbody=3, njnt=2, nq=2, nv=2, nu=1, na=0, ngeom=5, nsensordata=0)),
    ('cheetah', dict(nbody=8, njnt=9, nq=9, nv=9, nu=6, na=0, ngeom=9, nsensordata=3)),
    ('humanoid', dict(nbody=17, njnt=22, nq=28, nv=27, nu=21, na=0, ngeom=20, nsensordata=66)),
    ('quadruped', dict(nbody=18, njnt=17, nq=23, nv=22, nu=12, na=12, ngeom=20, nsensordata=36, ntendon=12, neq=4)),
    ('cmu_humanoid', dict(nbody=32, njnt=57, nq=63, nv=62, nu=56, na=0, nsensordata=25)),
])
def test_fixture_sizes_match_survey_counts(name, sizes):
  """SURVEY.md §8 counted these from the reference XMLs with a stdlib parser, independently of the compiler."""
  m = tm.load(name)
  for k, v in sizes.items():
    assert getattr(m, k) == v, (name, k, getattr(m, k), v)


def test_suite_model_parameters():
  """Spot values the reference files state: time steps and integrators, humanoid geom condim
  (`suite/humanoid.xml:13`), cheetah friction (`suite/cheetah.xml:11`), cartpole contacts disabled (`:7`)."""
  assert tm.load('cartpole').opt.integrator == 1 and tm.load('cartpole').opt.timestep == .01
  assert tm.load('cheetah').opt.timestep == .01 and tm.load('humanoid').opt.timestep == .005
  h = tm.load('humanoid')
  floor = h.name2id('floor', 'geom')
  assert h.geom_condim[floor] == 3 and all(h.geom_condim[g] == 1 for g in range(h.ngeom) if g != floor)
  c = tm.load('cheetah')
  assert all(abs(c.geom_friction[g][0] - .4) < 1e-15 for g in range(1, c.ngeom))
  assert tm.load('cartpole').opt.disableflags & (1 << 4)
