"""Run the reference's own dm_control/suite/suite_test.py, unmodified, over ALL of its tasks against this repo's engine
(CPU emulation build of the kernels), compiler and reference-facing Physics view; print the tally per failure reason.
Needs /root/reference. Takes tens of minutes (2 x 1000 control steps per task in one of the tests).
Usage: python tools/run_reference_suite_test.py [substring of test ids ...]
"""
import collections
import importlib
import os
import sys
import unittest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'emu')):
  sys.path.insert(0, p)
os.environ['B200MJ_EMULATE_GPU'] = '1'
import gpu_shim; gpu_shim.install()                                   # noqa: E402,E702
import refshim; refshim.install(); refshim.install_suite_package()   # noqa: E402,E702


def flatten(s):
  for t in s:
    if isinstance(t, unittest.TestSuite):
      yield from flatten(t)
    else:
      yield t


mod = importlib.import_module('dm_control.suite.suite_test')
tests = list(flatten(unittest.defaultTestLoader.loadTestsFromModule(mod)))
if len(sys.argv) > 1:
  tests = [t for t in tests if any(k in t.id() for k in sys.argv[1:])]
print(len(tests), 'tests selected')
res = unittest.TextTestRunner(verbosity=1).run(unittest.TestSuite(tests))
bad = collections.Counter()
per_task = collections.defaultdict(list)
for t, tb in res.failures + res.errors:
  reason = tb.strip().splitlines()[-1][:110]
  bad[reason] += 1
print('RESULT run', res.testsRun, 'passed', res.testsRun - len(res.failures) - len(res.errors), 'failures', len(res.failures), 'errors', len(res.errors))
for reason, n in bad.most_common():
  print('%4d  %s' % (n, reason))
