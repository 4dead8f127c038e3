"""PGS (`mj_solPGS`): BASELINE.json's north_star names "the PGS/Newton constraint-solver iterations"; the reference's testing
humanoid selects it (dm_control/mujoco/testing/assets/humanoid.xml:9). Oracle: the dual problem it solves must give the
same physics as the Newton solver's primal problem; CUDA path (fused kernel): parity with the oracle."""
import numpy as np
import pytest

from dm_control_b200 import mjcf_compile, testing_models as tm


def _pgs(xml, iterations=50, tolerance=None, **caps):
  opt = f'solver="PGS" iterations="{iterations}"' + (f' tolerance="{tolerance}"' if tolerance else '')
  if '<option' in xml:
    xml = xml.replace('<option', f'<option {opt}', 1)
  else:
    xml = xml.replace('<mujoco>', f'<mujoco><option {opt}/>', 1)
  return mjcf_compile.compile_xml(xml, **caps)


def test_pgs_reaches_the_newton_solution_at_rest(oracle_mod):
  """A free box on a plane (wrapper/core_test.py:393-416 geometry): both solvers must hold it with contact forces that
  sum to its weight; PGS with a generous budget lands on the Newton state."""
  mn = mjcf_compile.compile_xml(tm.XML['free_box'], nconmax=8, njmax=32)
  mp = _pgs(tm.XML['free_box'], iterations=500, tolerance=1e-12, nconmax=8, njmax=32)
  on, op = oracle_mod.OraclePhysics(mn), oracle_mod.OraclePhysics(mp)
  for o in (on, op):
    o.forward()
    for _ in range(500):
      o.control_step(1)
  assert op.ncon == on.ncon == 4
  weight = 9.81 * mp.body_mass[1]
  fp = sum(op.efc_force[c.efc_address:c.efc_address + 4].sum() for c in op.contact)
  assert abs(fp - weight) < 1e-6 * weight
  assert np.abs(op.qpos - on.qpos).max() < 1e-6
  assert (np.asarray(op.efc_force)[:op.nefc] >= 0).all()          # unilateral rows never pull


def test_pgs_tracks_newton_through_a_contact_rich_rollout(oracle_mod):
  mn = mjcf_compile.compile_xml(tm.XML['pendulum_free'], nconmax=16, njmax=64)
  mp = _pgs(tm.XML['pendulum_free'], iterations=100, nconmax=16, njmax=64)
  q0, v0 = tm.initial_states(mp, 'pendulum_free', 1, 3)
  on, op = oracle_mod.OraclePhysics(mn), oracle_mod.OraclePhysics(mp)
  for o in (on, op):
    o.qpos[:] = q0[0]; o.qvel[:] = v0[0]; o.forward()
  seen = 0
  for _ in range(60):
    on.control_step(2); op.control_step(2)
    seen += op.ncon
  assert seen > 20 and 1 <= op.solver_niter <= 100
  assert np.abs(op.qpos - on.qpos).max() < 5e-3                    # an iterative dual solver: close, not equal


@pytest.mark.gpu
@pytest.mark.parametrize('name,nsub,nsteps,caps', [('pendulum_free', 2, 30, dict(nconmax=16, njmax=64)),
                                                   ('free_box', 1, 200, dict(nconmax=8, njmax=32))])
def test_cuda_pgs_matches_oracle(name, nsub, nsteps, caps, oracle_mod):
  import torch
  from dm_control_b200.physics import BatchedPhysics
  m = _pgs(tm.XML[name], iterations=50, **caps)
  B = 3
  q0, v0 = tm.initial_states(m, name, B, 5)
  phys = BatchedPhysics(m, batch=B)
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  tape = np.random.RandomState(2).uniform(-1, 1, (nsteps, B, max(m.nu, 1)))
  for t in range(nsteps):
    if m.nu:
      phys.set_control(torch.as_tensor(tape[t]))
    phys.step(nsub)
  for e in range(B):
    o = oracle_mod.OraclePhysics(m)
    o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward()
    for t in range(nsteps):
      if m.nu:
        o.ctrl[:] = tape[t, e]
      o.control_step(nsub)
    assert np.abs(phys.data.qpos[e].cpu().numpy() - o.qpos).max() < 1e-7
    assert np.abs(phys.data.qvel[e].cpu().numpy() - o.qvel).max() < 1e-6
    assert int(phys.data.ncon[e]) == o.ncon and int(phys.data.solver_niter[e]) == o.solver_niter
