"""Physics-law pins of the oracle on the benchmark models themselves (CPU).

MuJoCo is absent, so no trajectory goldens exist. Besides the reference's analytic known-answers
(`test_oracle_kat.py`, `test_lqr.py`) the restatement is held to identities every rigid-body engine must satisfy,
each one tying together stages that are computed by *independent* code paths:

  * kinetic energy: v'Mv/2 (composite-rigid-body M)  ==  sum over bodies of m|v_c|^2/2 + w'Iw/2 (from `cvel`)
  * gravity:        qfrc_bias(q, 0) (recursive Newton-Euler)  ==  d(potential energy)/dq (from `xipos`)
  * Coriolis:       v'[bias(q, v) - bias(q, 0)]  ==  v'(dM/dt)v/2  (RNE against the derivative of CRB)
  * momentum:       with gravity and contacts off, actuator and damping forces are internal: d/dt of the total
                    linear and angular momentum along the solved acceleration is zero (pins actuator moments,
                    passive forces, M^-1 and the bias together)

The CUDA path is held to the oracle by the parity tests, so these pins carry over.
"""
import numpy as np
import pytest

from dm_control_b200 import testing_models as tm
from oracle import oracle as om

MODELS = ['cheetah', 'humanoid', 'quadruped', 'cmu_humanoid']
FREE, BALL, SLIDE, HINGE = 0, 1, 2, 3


def _quat_mul(a, b):
  w1, x1, y1, z1 = a; w2, x2, y2, z2 = b
  return np.array([w1*w2 - x1*x2 - y1*y2 - z1*z2, w1*x2 + x1*w2 + y1*z2 - z1*y2,
                   w1*y2 - x1*z2 + y1*w2 + z1*x2, w1*z2 + x1*y2 - y1*x2 + z1*w2])


def _quat_exp(w):
  a = np.linalg.norm(w)
  if a < 1e-300:
    return np.array([1.0, 0, 0, 0])
  return np.concatenate([[np.cos(a / 2)], np.sin(a / 2) * w / a])


def integrate(model, qpos, vel, h):
  """qpos advanced by h along the dof-space velocity `vel` (free/ball angular velocity is in the local frame)."""
  q = qpos.copy()
  for j in range(model.njnt):
    t, qa, da = int(model.jnt_type[j]), int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
    if t == FREE:
      q[qa:qa + 3] += h * vel[da:da + 3]
      q[qa + 3:qa + 7] = _quat_mul(q[qa + 3:qa + 7], _quat_exp(h * vel[da + 3:da + 6]))
    elif t == BALL:
      q[qa:qa + 4] = _quat_mul(q[qa:qa + 4], _quat_exp(h * vel[da:da + 3]))
    else:
      q[qa] += h * vel[da]
  return q


def random_state(model, seed, vel_scale=1.0):
  rs = np.random.RandomState(seed)
  q = np.asarray(model.qpos0, dtype=np.float64).copy()
  for j in range(model.njnt):
    t, qa = int(model.jnt_type[j]), int(model.jnt_qposadr[j])
    if t == FREE:
      q[qa:qa + 3] += rs.uniform(-.3, .3, 3) + np.array([0, 0, 1.5])
      quat = rs.randn(4); q[qa + 3:qa + 7] = quat / np.linalg.norm(quat)
    elif t == BALL:
      quat = rs.randn(4); q[qa:qa + 4] = quat / np.linalg.norm(quat)
    else:
      lo, hi = (model.jnt_range[j] if model.jnt_limited[j] else (-1.0, 1.0))
      q[qa] = rs.uniform(0.7 * lo + 0.3 * hi, 0.3 * lo + 0.7 * hi) if t == HINGE else rs.uniform(-.2, .2)
  return q, vel_scale * rs.randn(model.nv)


def set_state(o, q, v):
  o.qpos[:] = q; o.qvel[:] = v
  o.forward()


def body_com_velocity(model, o):
  """-> (angular [nb,3], linear-at-body-COM [nb,3]) from the subtree-COM based spatial velocities `cvel`."""
  cvel, xipos, scom = o.cvel, o.xipos, o.subtree_com
  root = np.asarray(model.body_rootid)
  ang = cvel[:, :3]
  lin = cvel[:, 3:] + np.cross(ang, xipos - scom[root])
  return ang, lin


def momenta(model, o):
  """Total linear momentum and angular momentum about the world origin."""
  ang, lin = body_com_velocity(model, o)
  mass = np.asarray(model.body_mass)
  P = (mass[:, None] * lin).sum(0)
  L = np.zeros(3)
  for b in range(1, model.nbody):
    R = o.ximat[b].reshape(3, 3)
    I = R @ np.diag(model.body_inertia[b]) @ R.T
    L += I @ ang[b] + mass[b] * np.cross(o.xipos[b], lin[b])
  return P, L


@pytest.mark.parametrize('name', MODELS)
def test_kinetic_energy_two_ways(name):
  model = tm.load(name)
  o = om.OraclePhysics(model)
  for seed in range(3):
    q, v = random_state(model, seed)
    set_state(o, q, v)
    ke_m = 0.5 * v @ o.M_dense() @ v - 0.5 * np.sum(np.asarray(model.dof_armature) * v * v)
    ang, lin = body_com_velocity(model, o)
    ke_b = 0.0
    for b in range(1, model.nbody):
      R = o.ximat[b].reshape(3, 3)
      wl = R.T @ ang[b]
      ke_b += 0.5 * model.body_mass[b] * lin[b] @ lin[b] + 0.5 * wl @ (np.asarray(model.body_inertia[b]) * wl)
    np.testing.assert_allclose(ke_m, ke_b, rtol=1e-10)


@pytest.mark.parametrize('name', MODELS)
def test_gravity_bias_is_potential_gradient(name):
  model = tm.load(name)
  o = om.OraclePhysics(model)
  g = np.asarray(model.opt.gravity, dtype=np.float64)
  mass = np.asarray(model.body_mass)
  def potential(qq):
    set_state(o, qq, np.zeros(model.nv))
    return -float((mass[:, None] * o.xipos).sum(0) @ g)
  q, _ = random_state(model, 11)
  set_state(o, q, np.zeros(model.nv))
  bias = o.qfrc_bias.copy()
  eps = 1e-6
  grad = np.zeros(model.nv)
  for i in range(model.nv):
    e = np.zeros(model.nv); e[i] = 1
    grad[i] = (potential(integrate(model, q, e, eps)) - potential(integrate(model, q, e, -eps))) / (2 * eps)
  np.testing.assert_allclose(bias, grad, rtol=1e-6, atol=1e-6 * np.abs(grad).max())


@pytest.mark.parametrize('name', MODELS)
def test_coriolis_power_identity(name):
  model = tm.load(name)
  o = om.OraclePhysics(model)
  q, v = random_state(model, 5)
  set_state(o, q, v); bias_v = o.qfrc_bias.copy()
  set_state(o, q, 0 * v); bias_0 = o.qfrc_bias.copy()
  eps = 1e-6
  set_state(o, integrate(model, q, v, eps), v); Mp = o.M_dense().copy()
  set_state(o, integrate(model, q, v, -eps), v); Mm = o.M_dense().copy()
  lhs = v @ (bias_v - bias_0)
  rhs = 0.5 * v @ ((Mp - Mm) / (2 * eps)) @ v
  scale = abs(v @ Mp @ v) * np.linalg.norm(v)
  assert abs(lhs - rhs) < 1e-6 * scale, (lhs, rhs)


@pytest.mark.parametrize('name', ['humanoid', 'quadruped', 'cmu_humanoid'])
def test_internal_forces_conserve_momentum(name):
  """Free-floating models only (the cheetah's root is pinned to the world by slide/hinge joints)."""
  model = tm.load(name).copy()
  o = om.OraclePhysics(model)
  o.disableflags = int(o.disableflags) | (1 << 4) | (1 << 6)          # contact, gravity off: every remaining force is internal
  rs = np.random.RandomState(2)
  q, v = random_state(model, 9, vel_scale=0.5)
  set_state(o, q, v)
  o.ctrl[:] = rs.uniform(-1, 1, model.nu)
  if model.na:
    o.act[:] = rs.uniform(-1, 1, model.na)
  o.forward()
  assert o.nefc == 0 or name == 'quadruped'      # joint limits are not hit in the sampled range; quadruped carries tendon equalities (internal)
  a = o.qacc.copy()
  ctrl = o.ctrl.copy()
  eps = 1e-6
  out = []
  for s in (+1, -1):
    set_state(o, integrate(model, q, v, s * eps), v + s * eps * a)
    out.append(momenta(model, o))
  dP = (out[0][0] - out[1][0]) / (2 * eps)
  dL = (out[0][1] - out[1][1]) / (2 * eps)
  set_state(o, q, v)
  P, L = momenta(model, o)
  mass = float(np.sum(model.body_mass))
  force_scale = np.abs(o.M_dense() @ a).max()
  assert np.abs(dP).max() < 1e-5 * max(force_scale, 1.0), dP
  assert np.abs(dL).max() < 1e-5 * max(force_scale, 1.0), dL


@pytest.mark.parametrize('name', ['humanoid', 'quadruped', 'cmu_humanoid'])
def test_total_momentum_rate_is_weight(name):
  """Newton's second law for the whole mechanism: with contacts off, dP/dt = (total mass) * gravity whatever the
  actuators do."""
  model = tm.load(name).copy()
  o = om.OraclePhysics(model)
  o.disableflags = int(o.disableflags) | (1 << 4)
  q, v = random_state(model, 21, vel_scale=0.5)
  set_state(o, q, v)
  o.ctrl[:] = np.random.RandomState(4).uniform(-1, 1, model.nu)
  o.forward()
  a = o.qacc.copy()
  eps, out = 1e-6, []
  for s in (+1, -1):
    set_state(o, integrate(model, q, v, s * eps), v + s * eps * a)
    out.append(momenta(model, o)[0])
  dP = (out[0] - out[1]) / (2 * eps)
  weight = float(np.sum(model.body_mass)) * np.asarray(model.opt.gravity, dtype=np.float64)
  np.testing.assert_allclose(dP, weight, rtol=1e-6, atol=1e-6 * np.abs(weight).max())


@pytest.mark.parametrize('theta_deg,mu,slides', [(20, 0.5, False), (10, 0.3, False), (35, 0.5, True), (30, 0.3, True)])
def test_coulomb_stick_slip_threshold_on_an_incline(theta_deg, mu, slides):
  """A flat box on a plane under tilted gravity sticks when tan(theta) < mu and slides when tan(theta) > mu, and while
  it slides the time-averaged friction is the Coulomb force mu*m*g*cos(theta) (pyramidal cone: exact along a pyramid
  axis; the soft contact chatters, hence the 5 % band on the average)."""
  import math
  from dm_control_b200 import mjcf_compile
  th, g, dt = math.radians(theta_deg), 9.81, 0.002
  xml = (f'<mujoco><option timestep="{dt}" gravity="{g * math.sin(th)} 0 {-g * math.cos(th)}"/><worldbody>'
         f'<geom name="floor" type="plane" size="50 50 .1" friction="{mu} .005 .0001"/>'
         f'<body name="box" pos="0 0 .02"><freejoint/><geom name="box" type="box" size=".3 .3 .02" friction="{mu} .005 .0001"/></body>'
         '</worldbody></mujoco>')
  o = om.OraclePhysics(mjcf_compile.compile_xml(xml))
  o.forward()
  v = []
  for _ in range(600):
    o.step(1)
    v.append(float(o.qvel[0]))
  if not slides:
    assert abs(v[-1]) < 5e-3 and abs(v[-1] - v[-201]) < 1e-6          # at rest (soft-constraint creep only)
  else:
    a = (v[-1] - v[-401]) / (400 * dt)
    friction = g * math.sin(th) - a
    assert friction == pytest.approx(mu * g * math.cos(th), rel=0.05)
