"""Which of the reference's suite tasks run UNMODIFIED on the engine? (this container only: needs /root/reference)

For every domain under /root/reference/dm_control/suite and every task in its SUITE: import the reference's own task
file (tests/refshim), build the environment (the file's own MJCF editing, compiled on the fly by this repo's compiler),
reset, take 10 random-action control steps on the B = 1 reference-facing view (kernels: CPU emulation build) and compare
the trajectory with the oracle (tests/test_reference_tasks.py: run_unmodified). Prints the tasks that pass with their worst
|engine - oracle|, and the ones that are refused with the reason (unsupported MuJoCo features are refused loudly, never
approximated).  Run:  python tools/probe_reference_suite.py
"""
import os, sys, importlib, traceback
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tests'); sys.path.insert(0, ROOT+'/tests/emu')
os.environ['B200MJ_EMULATE_GPU']='1'
import gpu_shim; gpu_shim.install()
import test_reference_tasks as t
import refshim; refshim.install()
doms = ['acrobot','ball_in_cup','cartpole','cheetah','finger','fish','hopper','humanoid','humanoid_CMU','manipulator','pendulum','point_mass','quadruped','reacher','stacker','swimmer','walker','lqr','dog']
ok, bad = [], []
for dom in doms:
  try:
    mod = importlib.import_module('dm_control.suite.' + dom)
    tasks = list(mod.SUITE.keys())
  except Exception as ex:
    bad.append((dom, '*', 'import: ' + repr(ex)[:120])); continue
  for task in tasks:
    try:
      r = t.run_unmodified(dom, task, 10)
      ok.append((dom, task, '%.1e' % r['worst']))
    except BaseException as ex:
      bad.append((dom, task, repr(ex)[:160]))
print('OK', len(ok)); [print('  ', x) for x in ok]
print('FAILED', len(bad)); [print('  ', x) for x in bad]
