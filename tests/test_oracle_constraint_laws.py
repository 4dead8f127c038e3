"""Constraint-side pins of the oracle on contact-rich states of the benchmark models (CPU).

The Newton solver minimises  (a - a0)' M (a - a0) / 2 + sum_i s_i(J_i a - aref_i)  (SURVEY.md §8a). Whatever path the
iterates took, the returned acceleration must satisfy the optimality conditions of that convex problem; and the
contacts it acts on must be geometrically what the narrow phase claims. Checked after rollouts that leave the
models lying on / walking over the floor:

  * stationarity:  M qacc - qfrc_smooth - J' efc_force = 0   (to the solver tolerance)
  * force law:     efc_force_i = -D_i (J_i qacc - aref_i) on active rows, 0 on inactive ones; unilateral rows
                   (limits, pyramidal contact edges) never pull: force >= 0, and force = 0 => J_i qacc - aref_i >= 0
  * qfrc_constraint = J' efc_force
  * contacts: frame rows orthonormal and right-handed; every contact lies within margin; for geom-plane contacts the
    normal is the plane's z axis and `dist` is the signed height of the contact point pair over the plane;
    pyramid edge rows are J_n +- mu J_t
"""
import numpy as np
import pytest

from dm_control_b200 import testing_models as tm
from oracle import oracle as om

EQUALITY, FRICTION_DOF, FRICTION_TENDON, LIMIT_JOINT, LIMIT_TENDON, CONTACT_FRICTIONLESS, CONTACT_PYRAMIDAL = range(7)
PLANE = 0


def settled(name, steps, seed=0):
  model = tm.load(name)
  q0, v0 = tm.initial_states(model, name, 1, seed)
  o = om.OraclePhysics(model)
  o.qpos[:] = q0[0]; o.qvel[:] = v0[0]; o.forward()
  rs = np.random.RandomState(seed)
  extra = 0
  for i in range(steps + 2000):
    o.ctrl[:] = rs.uniform(-1, 1, model.nu)
    o.step(1)
    extra += o.ncon > 0
    if i >= steps and extra >= 40:          # at least `steps` steps, the last 40 or more of them in contact
      break
  o.ctrl[:] = rs.uniform(-1, 1, model.nu)
  o.forward()
  return model, o


@pytest.mark.parametrize('name,steps', [('humanoid', 300), ('humanoid', 120), ('cheetah', 150), ('quadruped', 250)])
def test_newton_solution_satisfies_kkt(name, steps):
  model, o = settled(name, steps)
  n = o.nefc
  assert n > 0 and o.ncon > 0, 'state should be in contact'
  J, D, aref, f = o.efc('efc_J'), o.efc('efc_D'), o.efc('efc_aref'), o.efc('efc_force')
  etype = o.efc('efc_type')
  M, a = o.M_dense(), o.qacc
  jar = J @ a - aref
  # stationarity of the primal cost
  r = M @ a - o.qfrc_smooth - J.T @ f
  scale = 1.0 / (float(model.stat.meaninertia) * max(1, model.nv))
  assert np.linalg.norm(r) * scale < 1e-6, np.linalg.norm(r) * scale
  # force law and unilaterality
  uni = etype != EQUALITY
  active = ~uni | (jar < 0)
  np.testing.assert_allclose(f[active], -D[active] * jar[active], rtol=1e-9, atol=1e-12)
  assert np.all(f[~active] == 0)
  assert np.all(f[uni] >= 0)
  np.testing.assert_allclose(o.qfrc_constraint, J.T @ f, rtol=1e-9, atol=1e-9)
  assert np.all(D > 0)


@pytest.mark.parametrize('name,steps', [('humanoid', 300), ('cheetah', 150), ('quadruped', 250)])
def test_contact_geometry_and_pyramid_rows(name, steps):
  model, o = settled(name, steps)
  J = o.efc('efc_J')
  gtype = np.asarray(model.geom_type)
  seen_plane = 0
  for c in o.contact:
    R = c.frame.reshape(3, 3)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert np.linalg.det(R) > 0.999999
    assert c.dist < c.includemargin + 1e-12
    g1, g2 = c.geom1, c.geom2
    assert g1 != g2 and model.geom_bodyid[g1] != model.geom_bodyid[g2]
    if gtype[g1] == PLANE:
      seen_plane += 1
      pz = o.geom_xmat[g1].reshape(3, 3)[:, 2]
      np.testing.assert_allclose(R[0], pz, atol=1e-12)                 # normal points from the plane into the other geom
      # contact point sits midway between the surfaces: its height over the plane is dist / 2
      h = (c.pos - o.geom_xpos[g1]) @ pz
      np.testing.assert_allclose(h, 0.5 * c.dist, atol=1e-12)
    if c.efc_address >= 0 and c.dim == 3:
      a0 = c.efc_address
      mu = c.friction[:2]
      # rows: n + mu1 t1, n - mu1 t1, n + mu2 t2, n - mu2 t2  =>  pairwise sums are 2 n, differences 2 mu t
      n_row = 0.5 * (J[a0] + J[a0 + 1])
      np.testing.assert_allclose(0.5 * (J[a0 + 2] + J[a0 + 3]), n_row, atol=1e-12)
      t1, t2 = 0.5 * (J[a0] - J[a0 + 1]) / mu[0], 0.5 * (J[a0 + 2] - J[a0 + 3]) / mu[1]
      # the three rows are the contact-frame components of the relative point velocity Jacobian: check through qvel
      v = o.qvel
      b1, b2 = model.geom_bodyid[g1], model.geom_bodyid[g2]
      def point_vel(b):
        if b == 0:
          return np.zeros(3)
        root = model.body_rootid[b]
        cv = o.cvel[b]
        return cv[3:] + np.cross(cv[:3], c.pos - o.subtree_com[root])
      rel = point_vel(b2) - point_vel(b1)
      np.testing.assert_allclose([n_row @ v, t1 @ v, t2 @ v], R @ rel, atol=1e-9)
  assert seen_plane > 0
