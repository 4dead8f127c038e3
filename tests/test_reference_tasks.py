"""north_star: "suite tasks run unchanged". The UNMODIFIED reference sources — dm_control/rl/control.py (Environment),
dm_control/suite/{humanoid,cartpole,cheetah,quadruped}.py (tasks + Physics subclasses, incl. the model editing cartpole.py
and quadruped.py do with lxml), suite/base.py, suite/common, suite/utils/randomizers.py, utils/rewards.py,
utils/containers.py, utils/xml_tools.py — are imported from /root/reference (tests/refshim wires the few absent
third-party modules) and drive a B = 1 view of the batched CUDA engine (dm_control_b200/refview.py): humanoid:run for 100
control steps, the other BASELINE suite configs and eight further domains for 20 (tools/probe_reference_suite.py sweeps all 47
tasks of the reference suite: 30 run, 17 are refused for a named unsupported feature). The trajectory is checked against the CPU oracle stepped from the
same post-reset state with the same actions.

/root/reference exists in the build container only: the tests skip where it is absent (GPU box).
  * `-m gpu` twin: runs on the device (or under B200MJ_EMULATE_GPU=1);
  * CPU-collectable test: runs the same body in a child process against the CPU emulation build of the kernels."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refshim   # noqa: E402

needs_reference = pytest.mark.skipif(not refshim.available(), reason='/root/reference is not on this machine')


def run_unmodified_humanoid(nsteps=100):
  refshim.install()
  import dm_control.suite.humanoid as ref_humanoid            # the reference file itself
  assert os.path.realpath(ref_humanoid.__file__).startswith(os.path.realpath(refshim.REFERENCE))
  from oracle import oracle as om
  env = ref_humanoid.run(random=7)                            # control.Environment(Physics, Humanoid, ...), unmodified
  spec = env.action_spec()
  assert spec.shape == (21,) and spec.minimum.min() == -1 and spec.maximum.max() == 1
  ts = env.reset()
  assert ts.first() and ts.reward is None
  assert env.physics.data.ncon == 0                           # humanoid.py:160-166: re-drawn until contact-free
  assert set(ts.observation) == {'joint_angles', 'head_height', 'extremities', 'torso_vertical', 'com_velocity', 'velocity'}
  phys = env.physics
  o = om.OraclePhysics(phys.model)
  o.qpos[:] = phys.data.qpos; o.qvel[:] = phys.data.qvel; o.forward()
  rs = np.random.RandomState(3)
  worst = 0.0
  for t in range(nsteps):
    a = rs.uniform(-1, 1, 21)
    ts = env.step(a)
    o.ctrl[:] = a; o.control_step(5)                          # control_timestep .025 / timestep .005
    assert not ts.last() and ts.discount == 1.0
    assert 0.0 <= ts.reward <= 1.0
    worst = max(worst, float(np.abs(phys.data.qpos - o.qpos).max()), float(np.abs(phys.data.qvel - o.qvel).max()) * 0.1)
    assert phys.data.ncon == o.ncon
    assert [(c.geom1, c.geom2) for c in phys.data.contact] == [(c.geom1, c.geom2) for c in o.contact]
    # the reference task's own observation code, on the oracle's numbers
    np.testing.assert_allclose(ts.observation['joint_angles'], o.qpos[7:], atol=1e-6)
    np.testing.assert_allclose(ts.observation['head_height'], np.asarray(o.xpos).reshape(-1, 3)[phys.model.name2id('head', 'body'), 2], atol=1e-6)
  assert abs(env.physics.time() - nsteps * 0.025) < 1e-9
  assert worst < 1e-6, worst
  return worst


def run_unmodified(domain, task, nsteps=20, seed=5):
  """Any of the BASELINE suite configs through the reference's own task file: build (the file's own model editing where it
  has any: cartpole.py:104-127, quadruped.py:55-93), reset (its own randomisation), step with random actions; the
  trajectory against the oracle stepped from the same post-reset state. Convex (MPR) contacts end the comparison of an
  episode (discontinuous in the pose)."""
  import importlib
  refshim.install()
  mod = importlib.import_module('dm_control.suite.' + domain)
  assert os.path.realpath(mod.__file__).startswith(os.path.realpath(refshim.REFERENCE))
  from oracle import oracle as om
  env = getattr(mod, task)(random=seed)
  spec = env.action_spec()
  ts = env.reset()
  assert ts.first() and ts.reward is None
  phys = env.physics
  nsub = int(round(env.control_timestep() / phys.timestep()))
  o = om.OraclePhysics(phys.model)
  o.qpos[:] = phys.data.qpos; o.qvel[:] = phys.data.qvel
  if phys.model.na:
    o.act[:] = phys.data.act
  o.forward()
  gtype = np.asarray(phys.model.geom_type)
  rs = np.random.RandomState(11)
  worst, compared = 0.0, 0
  for t in range(nsteps):
    a = rs.uniform(spec.minimum, spec.maximum)
    ts = env.step(a)
    o.ctrl[:] = a; o.control_step(nsub)
    assert ts.reward is not None and 0.0 <= ts.reward <= 1.0 + 1e-12
    if any(gtype[c.geom1] != 0 and (gtype[c.geom1] > 3 or gtype[c.geom2] > 3) for c in o.contact):
      break
    worst = max(worst, float(np.abs(phys.data.qpos - o.qpos).max()), float(np.abs(phys.data.qvel - o.qvel).max()) * 0.1)
    assert phys.data.ncon == o.ncon
    assert [(c.geom1, c.geom2) for c in phys.data.contact] == [(c.geom1, c.geom2) for c in o.contact]
    compared += 1
  assert abs(env.physics.time() - (t + 1) * env.control_timestep()) < 1e-9
  assert compared >= min(10, nsteps) and worst < 1e-6, (compared, worst)
  return dict(worst=worst, compared=compared, nsub=nsub, obs=sorted(ts.observation))


_SUITE_CASES = [
    ('cartpole', 'swingup', ['position', 'velocity']),                      # BASELINE.json config 0
    ('cartpole', 'balance', ['position', 'velocity']),
    ('cheetah', 'run', ['position', 'velocity']),                           # config 1
    ('quadruped', 'walk', ['egocentric_state', 'force_torque', 'imu', 'torso_upright', 'torso_velocity']),      # config 3
    ('humanoid', 'stand', None), ('humanoid', 'walk', None),
    # domains with no batched twin here: the reference's file + this repo's compiler and engine are all there is
    ('humanoid_CMU', 'stand', None),      # the 62-dof CMU model
    ('acrobot', 'swingup', None), ('fish', 'upright', None), ('hopper', 'hop', None), ('pendulum', 'swingup', None),
    ('point_mass', 'hard', None),         # writes physics.model.wrap_prm in place (point_mass.py:101-112)
    ('reacher', 'hard', None), ('walker', 'run', None), ('cartpole', 'three_poles', None),
]


@pytest.fixture(scope='module')
def emulated_suite_results():
  """ONE child process (CPU emulation build of the kernels) runs every case; the parametrized tests read their entry."""
  import json
  cases = [(d, t) for d, t, _ in _SUITE_CASES]
  code = ("import os, sys, json, traceback; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r);"
          "import gpu_shim; gpu_shim.install();"
          "import test_reference_tasks as t\n"
          "out = {}\n"
          "for d, k in %r:\n"
          "  try: out[d + ':' + k] = t.run_unmodified(d, k)\n"
          "  except BaseException as ex: out[d + ':' + k] = dict(error=traceback.format_exc()[-1500:])\n"
          "print('RESULT', json.dumps(out))") % (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'emu'), cases)
  env = dict(os.environ, B200MJ_EMULATE_GPU='1')
  r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=1500)
  assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
  return json.loads(r.stdout.split('RESULT', 1)[1])


@needs_reference
@pytest.mark.parametrize('domain,task,keys', _SUITE_CASES)
def test_unmodified_reference_suite_tasks_under_emulation(domain, task, keys, emulated_suite_results):
  """The reference's own suite/<domain>.py drives the engine (CPU emulation build of the kernels, child process)."""
  out = emulated_suite_results[f'{domain}:{task}']
  assert 'error' not in out, out.get('error')
  assert out['compared'] >= 10 and out['worst'] < 1e-6, out
  if keys is not None:
    assert out['obs'] == sorted(keys), out


@needs_reference
@pytest.mark.gpu
def test_unmodified_reference_humanoid_task_on_the_engine():
  run_unmodified_humanoid(100)


@needs_reference
def test_unmodified_reference_humanoid_task_under_emulation():
  """CPU-collectable twin: the same body in a child process, kernels from the CPU emulation build (tests/emu)."""
  code = ("import os, sys; sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r);"
          "import gpu_shim; gpu_shim.install();"
          "import test_reference_tasks as t; print('worst', t.run_unmodified_humanoid(100))") % (
              ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'emu'))
  env = dict(os.environ, B200MJ_EMULATE_GPU='1')
  r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=900)
  assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
  assert 'worst' in r.stdout
