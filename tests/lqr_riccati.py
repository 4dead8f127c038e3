"""Closed-form pin for the LQR domain: cost-to-go Hessian, feedback gain and closed-loop decay rate of the discrete-time
LQR problem the domain poses (what dm_control/suite/lqr_solver.py:27-82 computes for the reference's own test,
suite/lqr_test.py:33-59). Test infrastructure: the product package does not ship it.

State x = (q, v) of a damped spring chain with joint-space inertia M, stiffness diag(k), damping diag(d), stepped by
semi-implicit Euler:  v' = v + dt M^-1 (-k q - d v + B u),  q' = q + dt v'  with the first `n_controls` joints actuated,
stage cost |q|^2 + c |u|^2."""
import numpy as np
import scipy.linalg


def riccati(mass, stiffness, damping, dt, n_controls, control_cost_coef):
  n = mass.shape[0]
  minv = np.linalg.inv(mass)
  accel_q, accel_v = -minv @ np.diag(stiffness), -minv @ np.diag(damping)      # dv/dt = accel_q q + accel_v v + M^-1 B u
  vel_next = np.hstack([dt * accel_q, np.eye(n) + dt * accel_v])               # v' as a function of (q, v)
  pos_next = np.hstack([np.eye(n), np.zeros((n, n))]) + dt * vel_next          # q' = q + dt v'
  a = np.vstack([pos_next, vel_next])
  sel = np.zeros((n, n_controls)); sel[:n_controls] = np.eye(n_controls)
  bv = dt * minv @ sel
  b = np.vstack([dt * bv, bv])
  q = np.diag(np.r_[np.ones(n), np.zeros(n)])
  r = control_cost_coef * np.eye(n_controls)
  p = scipy.linalg.solve_discrete_are(a, b, q, r)
  gain = -np.linalg.solve(r + b.T @ p @ b, b.T @ p @ a)
  rate = np.abs(np.linalg.eigvals(a + b @ gain)).max()
  if rate >= 1.0:
    raise RuntimeError('Controlled system is unstable.')
  return p, gain, rate


def riccati_for_env(env):
  """The same for a batched LQR environment (model, hence P, K, decay rate, shared by the batch)."""
  phys = env.physics
  model = phys.model
  phys.forward()
  mass = phys.data.qM[0].reshape(model.nv, model.nv).cpu().numpy()
  return riccati(mass, np.asarray(model.jnt_stiffness).ravel(), np.asarray(model.dof_damping).ravel(),
                 float(model.opt.timestep), model.nu, env.task.control_cost_coef)
