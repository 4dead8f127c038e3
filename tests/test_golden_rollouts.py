"""Seeded rollouts against committed golden files.

Two kinds of golden file can sit under tests/golden/:
  * oracle_rollouts.npz  — written by tools/make_golden_rollouts.py from the CPU oracle itself: a drift alarm for the
    oracle / the MJCF compiler, and the fixture the CUDA path is held to on the GPU box (where the oracle also runs, but
    the file additionally pins "what the oracle said on the day it was committed");
  * mujoco_rollouts.npz  — written by tools/dump_mujoco_goldens.py from the REAL `mujoco.mj_step` on a machine that
    has MuJoCo (absent from this image). When present, every test below ALSO checks against it at the tolerance
    BASELINE.json's north_star states (1e-5 relative on qpos / qvel, contact pairs exact) and reports it in its id.

Checked per case: final qpos / qvel / sensordata, the ncon trace of every control step, and the final contact list
(geom1, geom2) of every environment in order — "contact-pair indexing bit-exact".
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import make_golden_rollouts as mg   # noqa: E402

GOLD = np.load(mg.OUT)
_MJ = os.path.join(os.path.dirname(mg.OUT), 'mujoco_rollouts.npz')
MJGOLD = np.load(_MJ) if os.path.exists(_MJ) else None
KIND = 'oracle+mujoco' if MJGOLD is not None else 'oracle-only'


def test_golden_kind_is_reported():
  """Which external anchor the parity tests below have: none ('oracle-only': parity unpinned) or real mj_step dumps."""
  print(f'golden kind: {KIND}')
  assert KIND in ('oracle-only', 'oracle+mujoco')


def _split(pairs, ncon):
  out, at = [], 0
  for n in ncon:
    out.append(pairs[at:at + n]); at += n
  return out


def _check(name, got, qtol, vtol, stol, strict_after_mpr=True):
  """`strict_after_mpr=False` (the CUDA path): an environment is held to the golden trajectory step by step up to the
  first control step with an MPR contact (GOLD mpr_step: those contacts are discontinuous in the pose, so fused
  multiply-add rounding may take the GPU elsewhere afterwards); final state, contact list and sensors are compared for
  the environments that never had one. The oracle itself must reproduce everything."""
  mpr = GOLD[f'{name}_mpr_step']
  nsteps = GOLD[f'{name}_ncon_trace'].shape[0]
  for e in range(len(mpr)):
    upto = nsteps if strict_after_mpr else int(mpr[e])
    np.testing.assert_allclose(got['qpos_trace'][:upto, e], GOLD[f'{name}_qpos_trace'][:upto, e], rtol=0, atol=qtol)
    np.testing.assert_array_equal(got['ncon_trace'][:upto + (0 if strict_after_mpr else 1), e][:nsteps],
                                  GOLD[f'{name}_ncon_trace'][:upto + (0 if strict_after_mpr else 1), e][:nsteps])
  full = np.ones(len(mpr), bool) if strict_after_mpr else (mpr >= nsteps)
  np.testing.assert_allclose(got['qpos'][full], GOLD[f'{name}_qpos'][full], rtol=0, atol=qtol)
  np.testing.assert_allclose(got['qvel'][full], GOLD[f'{name}_qvel'][full], rtol=0, atol=vtol)
  gp, wp = _split(got['pairs'], got['ncon']), _split(GOLD[f'{name}_pairs'], GOLD[f'{name}_ncon'])
  for e in np.nonzero(full)[0]:
    np.testing.assert_array_equal(gp[e], wp[e])
  g = GOLD[f'{name}_sensordata']
  np.testing.assert_allclose(got['sensordata'][full], g[full], rtol=0, atol=stol * (1 + np.abs(g).max() if g.size else 1))
  if MJGOLD is not None and f'{name}_qpos' in MJGOLD.files:
    # real mj_step: north_star tolerance (<= 1e-5 relative after the fixed horizon), contact pairs exact
    for f in ('qpos', 'qvel'):
      ref = MJGOLD[f'{name}_{f}']
      assert np.abs(got[f] - ref).max() <= 1e-5 * (1 + np.abs(ref).max()), (name, f)
    np.testing.assert_array_equal(got['ncon_trace'], MJGOLD[f'{name}_ncon_trace'])
    np.testing.assert_array_equal(got['pairs'], MJGOLD[f'{name}_pairs'])


@pytest.mark.parametrize('name,nsub,nsteps', mg.CASES)
def test_oracle_reproduces_goldens(name, nsub, nsteps, oracle_mod):
  # same compiler flags on the same ISA reproduce bit for bit; allow 1e-9 for a different host libm
  _check(name, mg.rollout(name, nsub, nsteps), 1e-9, 1e-8, 1e-8)


def test_goldens_end_in_contact():
  """The contact-pair assertion must bite: every contact-capable model ends with contacts in at least one environment."""
  for name in ('cheetah', 'humanoid', 'quadruped_floor', 'pendulum_free', 'convex_zoo_floor', 'cmu_humanoid'):
    assert len(GOLD[f'{name}_pairs']) > 0, name


def cuda_rollout(name, nsub, nsteps):
  import torch
  from dm_control_b200.physics import BatchedPhysics
  model, q0, v0, tape = mg.inputs(name, nsteps)
  phys = BatchedPhysics(model, batch=mg.B)
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  trace = np.zeros((nsteps, mg.B), np.int32)
  qtrace = np.zeros((nsteps, mg.B, model.nq))
  for t in range(nsteps):
    phys.set_control(torch.as_tensor(tape[t])); phys.step(nsub)
    trace[t] = phys.data.ncon.cpu().numpy()
    qtrace[t] = phys.data.qpos.cpu().numpy()
  cg = phys.data.contact_geom.cpu().numpy().reshape(mg.B, -1, 2)
  pairs = [cg[e, :trace[-1, e]] for e in range(mg.B)]
  return dict(qpos=phys.data.qpos.cpu().numpy(), qvel=phys.data.qvel.cpu().numpy(), sensordata=phys.data.sensordata.cpu().numpy(),
              ncon=trace[-1].copy(), ncon_trace=trace, pairs=mg.pack_pairs(pairs), qpos_trace=qtrace)


@pytest.mark.gpu
@pytest.mark.parametrize('name,nsub,nsteps', mg.CASES)
def test_cuda_reproduces_goldens(name, nsub, nsteps):
  _check(name, cuda_rollout(name, nsub, nsteps), 1e-7, 1e-6, 1e-6, strict_after_mpr=False)
