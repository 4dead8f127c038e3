"""The reference's LQR physics test (dm_control/suite/lqr_test.py:33-59) on the oracle (CPU) and on the CUDA path.

Under the optimal linear policy u = K x the accumulated cost must equal the Riccati cost-to-go x0' P x0 / 2 to 1e-3.
P and K come from the joint-space inertia, the joint stiffness and the time step alone (suite/lqr_solver.py:27-82),
so the test pins `M` of a serial chain, the passive spring force and the semi-implicit Euler update against a
closed form that needs no MuJoCo binary. As in the reference, the first loop iteration contributes 1 = |q0|^2 / 2
(initial positions lie on the sphere of radius sqrt 2, lqr.py:234-238).
"""
import math

import numpy as np

import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lqr_riccati   # noqa: E402
import pytest

from dm_control_b200.suite import lqr

TOL = 1e-3
LEVELS = [(2, 1), (6, 2)]


def _n_steps(beta):
  return int(math.ceil(math.log10(TOL) / math.log10(beta)))


@pytest.mark.parametrize('n_bodies,n_actuators', LEVELS)
def test_lqr_optimal_policy_oracle(n_bodies, n_actuators):
  from oracle import oracle as om
  model = lqr.compile_model(n_bodies, n_actuators, 0)
  assert (model.nq, model.nv, model.nu) == (n_bodies, n_bodies, n_actuators)
  o = om.OraclePhysics(model)
  o.forward()
  mass = o.M_dense()
  # serial chain of equal spheres on collinear sliders: M[i][j] = m * (n - max(i, j))
  m = 1000 * 4 / 3 * math.pi * 0.1 ** 3
  expect = m * (n_bodies - np.maximum.outer(np.arange(n_bodies), np.arange(n_bodies)))
  np.testing.assert_allclose(mass, expect, rtol=1e-12)
  p, k, beta = lqr_riccati.riccati(mass, np.asarray(model.jnt_stiffness).ravel(), np.asarray(model.dof_damping).ravel(),
                         float(model.opt.timestep), model.nu, 0.1)
  rs = np.random.RandomState(3)
  unit = rs.randn(n_bodies)
  o.qpos[:] = math.sqrt(2) * unit / np.linalg.norm(unit)
  o.qvel[:] = 0
  o.forward()
  x0 = np.hstack([o.qpos, o.qvel])
  total, reward = 0.0, None
  for _ in range(_n_steps(beta)):
    x = np.hstack([o.qpos, o.qvel])
    u = k.dot(x)
    total += 1 - (reward or 0.0)
    o.ctrl[:] = u
    o.control_step(1)
    reward = 1 - (0.5 * o.qpos.dot(o.qpos) + 0.5 * 0.1 * u.dot(u))
  np.testing.assert_allclose(0.5 * x0.dot(p).dot(x0), total, rtol=TOL)


def test_model_matches_reference_draw_order():
  """lqr.py:104-134 draws stiffness then damping per body from the task's RandomState."""
  model = lqr.compile_model(3, 2, 0)
  draws = np.random.RandomState(0).uniform(size=6)
  np.testing.assert_allclose(np.asarray(model.jnt_stiffness).ravel(), 15 + 10 * draws[0::2], rtol=1e-15)
  assert float(model.opt.timestep) == 0.03 and int(model.opt.disableflags) & 1
  with pytest.raises(ValueError, match='At most 1 actuator per body'):
    lqr.make_model_xml(1, 2, np.random.RandomState(0))
  with pytest.raises(ValueError, match='At least 1 body'):
    lqr.make_model_xml(0, 0, np.random.RandomState(0))


@pytest.mark.gpu
@pytest.mark.parametrize('level', ['lqr_2_1', 'lqr_6_2'])
def test_lqr_optimal_policy_gpu(level):
  import torch
  from dm_control_b200 import control, suite
  B = 32
  env = suite.load('lqr', level, batch=B, seed=0)
  p, k, beta = lqr_riccati.riccati_for_env(env)
  K = torch.as_tensor(k, device=env.physics.device)
  ts = env.reset()
  x0 = torch.cat([ts.observation['position'], ts.observation['velocity']], dim=1)
  assert torch.allclose(x0[:, :env.physics.model.nq].norm(dim=1), torch.full((B,), math.sqrt(2), dtype=torch.float64, device=x0.device))
  total = torch.zeros(B, dtype=torch.float64, device=x0.device)
  for i in range(_n_steps(beta)):
    x = torch.cat([ts.observation['position'], ts.observation['velocity']], dim=1)
    u = x @ K.T
    total += 1 - (ts.reward if ts.reward is not None else 0.0)
    ts = env.step(u)
    assert bool((ts.step_type == control.MID).all()) and bool((ts.discount == 1).all())
  P = torch.as_tensor(p, device=x0.device)
  expected = 0.5 * torch.einsum('bi,ij,bj->b', x0, P, x0)
  np.testing.assert_allclose(total.cpu().numpy(), expected.cpu().numpy(), rtol=TOL)


@pytest.mark.gpu
def test_lqr_termination_discount_zero():
  """lqr.py:262-267: the episode ends with discount 0 once the state norm is below 1e-6; the next step re-initialises."""
  import torch
  from dm_control_b200 import control, suite
  env = suite.load('lqr', 'lqr_2_1', batch=4, seed=1)
  env.reset()
  d = env.physics.data
  d.qpos[1:3] = 0; d.qvel[1:3] = 0                      # two environments already at the origin
  ts = env.step(torch.zeros(4, 1, dtype=torch.float64, device=env.physics.device))
  assert ts.step_type.tolist() == [control.MID, control.LAST, control.LAST, control.MID]
  assert ts.discount.tolist() == [1.0, 0.0, 0.0, 1.0]
  ts = env.step(torch.zeros(4, 1, dtype=torch.float64, device=env.physics.device))
  assert ts.step_type.tolist() == [control.MID] * 4
  assert torch.allclose(d.qpos[1:3].norm(dim=1), torch.full((2,), math.sqrt(2), dtype=torch.float64, device=d.qpos.device), atol=0.1)
