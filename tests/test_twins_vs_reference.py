"""The batched torch twins of reference utilities against the REFERENCE'S OWN implementation, imported unmodified from
/root/reference (tests/refshim) — not against a restatement. Skipped where the reference checkout is absent (GPU box);
the restatement-based tests (tests/test_rewards_and_env_cpu.py, tests/test_randomizers_cpu.py) run everywhere.

  * dm_control_b200/rewards.py            vs dm_control/utils/rewards.py:25-135            (every sigmoid, bounds, margins)
  * dm_control_b200/control.compute_n_steps vs dm_control/rl/control.py:168-194            (values and error cases)
(The reference's randomizers and task files themselves run unmodified on the engine in tests/test_reference_tasks.py.)
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refshim   # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason='/root/reference is not on this machine')

SIGMOIDS = ('gaussian', 'hyperbolic', 'long_tail', 'reciprocal', 'cosine', 'linear', 'quadratic', 'tanh_squared')


@pytest.fixture(scope='module')
def ref():
  refshim.install()
  import dm_control.utils.rewards as ref_rewards
  import dm_control.rl.control as ref_control
  assert os.path.realpath(ref_rewards.__file__).startswith(os.path.realpath(refshim.REFERENCE))
  return ref_rewards, ref_control


@pytest.mark.parametrize('sigmoid', SIGMOIDS)
def test_tolerance_equals_the_reference(ref, sigmoid):
  from dm_control_b200 import rewards
  ref_rewards, _ = ref
  rs = np.random.RandomState(0)
  x = np.concatenate([rs.uniform(-6, 6, 400), [-1.0, 0.0, 0.5, 1.0, 2.0, 3.0]])
  for bounds, margin, vam in (((0.0, 0.0), 1.0, 0.1), ((-1.0, 2.0), 0.5, 0.3), ((1.4, float('inf')), 0.35, 0.1), ((3.0, 3.0), 3.0, 0.0 if sigmoid in ('cosine', 'linear', 'quadratic') else 0.05),
                              ((0.0, 1.0), 0.0, 0.1)):
    got = rewards.tolerance(torch.as_tensor(x), bounds=bounds, margin=margin, sigmoid=sigmoid, value_at_margin=vam).numpy()
    want = ref_rewards.tolerance(x, bounds=bounds, margin=margin, sigmoid=sigmoid, value_at_margin=vam)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-15)


def test_tolerance_errors_equal_the_reference(ref):
  from dm_control_b200 import rewards
  ref_rewards, _ = ref
  for kw in (dict(bounds=(1.0, 0.0)), dict(margin=-1.0), dict(margin=1.0, sigmoid='nope'), dict(margin=1.0, sigmoid='gaussian', value_at_margin=0.0),
             dict(margin=1.0, sigmoid='linear', value_at_margin=1.0)):
    with pytest.raises(ValueError) as a:
      ref_rewards.tolerance(np.array([0.5]), **kw)
    with pytest.raises(ValueError) as b:
      rewards.tolerance(torch.tensor([0.5], dtype=torch.float64), **kw)
    assert str(a.value) == str(b.value)


def test_compute_n_steps_equals_the_reference(ref):
  from dm_control_b200 import control
  _, ref_control = ref
  for ct, pt in ((0.025, 0.005), (0.03, 0.005), (0.02, 0.0025), (0.01, 0.01), (0.04, 0.002)):
    assert control.compute_n_steps(ct, pt) == ref_control.compute_n_steps(ct, pt)
  for ct, pt in ((0.005, 0.01), (0.0251, 0.005)):
    with pytest.raises(ValueError) as a:
      ref_control.compute_n_steps(ct, pt)
    with pytest.raises(ValueError) as b:
      control.compute_n_steps(ct, pt)
    assert str(a.value) == str(b.value)
