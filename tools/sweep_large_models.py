"""Offline sweep (no GPU): the kernel source on the CPU emulation against the oracle on the large-model paths — dual
Newton form, blocked 62 x 62 factorisation, collision queue — physics step by physics step over many seeds, ending an
environment's comparison at its first convex (MPR) contact (discontinuous in the pose). Prints every
deviation above 1e-8 (none over seeds 0-23 of cmu_humanoid, 10 control steps x 6 substeps, B = 4).
Usage: python tools/sweep_large_models.py cmu_humanoid 0 24 10 6      [B200MJ_TN=0 B200MJ_DUAL_MIN_NV=1 ... humanoid 0 16 16 5]
"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import b200mj_emu as emu
from dm_control_b200 import testing_models as tm
from oracle import oracle as om
om.build()
name = sys.argv[1]; seeds = range(int(sys.argv[2]), int(sys.argv[3])); nc = int(sys.argv[4]); nsub = int(sys.argv[5])
model = tm.load(name); gtype = np.asarray(model.geom_type)
B = 4
for seed in seeds:
  q0, v0 = tm.initial_states(model, name, B, seed)
  p = emu.EmuPhysics(model, B); p.data.qpos[:] = q0; p.data.qvel[:] = v0; p.forward()
  oracles = []
  for e in range(B):
    o = om.OraclePhysics(model); o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward(); oracles.append(o)
  tape = np.random.RandomState(seed + 100).uniform(-1, 1, (nc, B, model.nu))
  taint = [False] * B
  for t in range(nc):
    p.data.ctrl[:] = tape[t]
    for s in range(nsub):        # substep by substep so that convex contacts inside a control step are seen
      p.step(1)
      for e, o in enumerate(oracles):
        o.ctrl[:] = tape[t, e]; o.control_step(1)
        if not taint[e] and any(gtype[c.geom1] != 0 and (gtype[c.geom1] > 3 or gtype[c.geom2] > 3) for c in o.contact):
          taint[e] = True
        if taint[e]: continue
        d = max(np.abs(p.data.qpos[e] - o.qpos).max(), np.abs(p.data.qvel[e] - o.qvel).max() / 10)
        if d > 1e-8: print('seed', seed, 'step', t, 'sub', s, 'env', e, 'dev %.2e' % d, 'ncon', int(p.data.ncon[e]), o.ncon, 'nefc', int(p.data.nefc[e]), [(c.geom1, c.geom2, int(gtype[c.geom1]), int(gtype[c.geom2])) for c in o.contact]); taint[e] = True
print('done')
