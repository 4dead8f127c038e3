"""The reference's OWN test files, unmodified, against this repo's engine, compiler and reference-facing Physics view.

  * dm_control/suite/utils/randomizers_test.py — `randomize_limited_and_rotational_joints` on models the test compiles
    from inline MJCF (every joint type, limited ball joints, a singular inertia matrix): 5 tests;
  * dm_control/suite/loader_test.py — `suite.load` and the task-tag constants: 3 tests;
  * dm_control/suite/suite_test.py — the suite's conformance tests (named components, >= 2 cameras, observations /
    rewards / discounts conform to the specs, determinism, reward visualisation through material colours, environment
    kwargs, observations do not share memory / hold no constant elements over 2 x 1000 steps, randomised initial state),
    parameterised by the reference over its 51 tasks. Here: every test of the tasks listed in `_SUITE_TEST_TASKS`
    (tools/run_reference_suite_test.py runs the whole file; DESIGN.md 3 has the tally).

The files are imported from /root/reference through tests/refshim (absent third-party modules: dm_env, lxml, mock,
mujoco's enums / constants) and run in a child process on the CPU emulation build of the kernels. Skipped where the
reference checkout is absent (GPU box). They found one conformance bug when first run (accessors returning live arrays
instead of copies, engine.py:589-614) — fixed.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import refshim   # noqa: E402

pytestmark = pytest.mark.skipif(not refshim.available(), reason='/root/reference is not on this machine')

# (domain, task, with the 2 x 1000-step constant-elements test): the four BASELINE suite configs; the long test for cheetah only (time)
_SUITE_TEST_TASKS = [('cartpole', 'swingup', False), ('cheetah', 'run', True), ('humanoid', 'run', False), ('quadruped', 'walk', False)]

_CHILD = r'''
import os, sys, json, unittest, importlib
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests'); sys.path.insert(0, %(root)r + '/tests/emu')
import gpu_shim; gpu_shim.install()
import refshim; refshim.install(); refshim.install_suite_package()
mod = importlib.import_module(%(module)r)
assert os.path.realpath(mod.__file__).startswith(os.path.realpath(refshim.REFERENCE))
def flatten(s):
  for t in s:
    if isinstance(t, unittest.TestSuite): yield from flatten(t)
    else: yield t
tests = list(flatten(unittest.defaultTestLoader.loadTestsFromModule(mod)))
select = %(select)r
if select is not None:
  keep = []
  for t in tests:
    for dom, task, heavy in select:
      tag = "domain=%%r, task=%%r" %% (dom, task)
      tag2 = "(%%r, %%r)" %% (dom, task)
      if (tag in t.id() or tag2 in t.id()) and (heavy or 'constant_elements' not in t.id()):
        keep.append(t)
  tests = keep
res = unittest.TextTestRunner(verbosity=0, stream=open(os.devnull, 'w')).run(unittest.TestSuite(tests))
bad = [(t.id().split('.')[-1], tb.strip().splitlines()[-1][:200]) for t, tb in res.failures + res.errors]
print('RESULT', json.dumps(dict(run=res.testsRun, bad=bad)))
'''


def _run(module, select):
  code = _CHILD % dict(root=ROOT, module=module, select=select)
  r = subprocess.run([sys.executable, '-c', code], env=dict(os.environ, B200MJ_EMULATE_GPU='1'), capture_output=True, text=True, timeout=1700)
  assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
  return json.loads(r.stdout.split('RESULT', 1)[1])


def test_reference_randomizers_test_passes_unmodified():
  out = _run('dm_control.suite.utils.randomizers_test', None)
  assert out['run'] == 5 and out['bad'] == [], out


def test_reference_loader_test_passes_unmodified():
  out = _run('dm_control.suite.loader_test', None)      # suite.load with and without task kwargs, the tag constants
  assert out['run'] == 3 and out['bad'] == [], out


@pytest.mark.timeout(1800)
def test_reference_suite_test_passes_unmodified_for_the_baseline_configs():
  out = _run('dm_control.suite.suite_test', _SUITE_TEST_TASKS)
  # 8 parameterised tests per task (+ reward visualisation where the reference lists the task), minus the long one for three of them
  assert out['run'] >= 29 and out['bad'] == [], out
