"""CPU: the oracle (and the MJCF compiler behind the model fixtures) still reproduce the committed regression goldens.

tests/golden/oracle_rollouts.npz was written by tools/make_golden_rollouts.py from the oracle itself; it is a drift
alarm, not an external pin (DESIGN.md §3 lists the external pins). GPU: the CUDA path reproduces the same file.
"""
import os

import numpy as np
import pytest

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
import make_golden_rollouts as mg   # noqa: E402

GOLD = np.load(mg.OUT)


@pytest.mark.parametrize('name,nsub,nsteps', mg.CASES)
def test_oracle_reproduces_goldens(name, nsub, nsteps, oracle_mod):
  q, v, n, pairs = mg.rollout(name, nsub, nsteps)
  # same compiler flags on the same ISA reproduce bit for bit; allow 1e-9 for a different host libm
  np.testing.assert_allclose(q, GOLD[f'{name}_qpos'], rtol=0, atol=1e-9)
  np.testing.assert_allclose(v, GOLD[f'{name}_qvel'], rtol=0, atol=1e-8)
  np.testing.assert_array_equal(n, GOLD[f'{name}_ncon'])
  got = np.concatenate(pairs) if sum(len(p) for p in pairs) else np.zeros((0, 2), np.int32)
  np.testing.assert_array_equal(got, GOLD[f'{name}_pairs'])


@pytest.mark.gpu
@pytest.mark.parametrize('name,nsub,nsteps', mg.CASES)
def test_cuda_reproduces_goldens(name, nsub, nsteps):
  import torch
  from dm_control_b200 import testing_models as tm
  from dm_control_b200.physics import BatchedPhysics
  model = tm.load(name)
  B, seed = 3, 21
  q0, v0 = tm.initial_states(model, name, B, seed)
  tape = np.random.RandomState(seed + 1).uniform(-1, 1, (nsteps, B, model.nu))
  phys = BatchedPhysics(model, batch=B)
  phys.data.qpos.copy_(torch.as_tensor(q0)); phys.data.qvel.copy_(torch.as_tensor(v0)); phys.forward()
  for t in range(nsteps):
    phys.set_control(torch.as_tensor(tape[t])); phys.step(nsub)
  np.testing.assert_allclose(phys.data.qpos.cpu().numpy(), GOLD[f'{name}_qpos'], rtol=0, atol=1e-7)
  np.testing.assert_allclose(phys.data.qvel.cpu().numpy(), GOLD[f'{name}_qvel'], rtol=0, atol=1e-6)
  np.testing.assert_array_equal(phys.data.ncon.cpu().numpy(), GOLD[f'{name}_ncon'])
