"""Write tests/golden/mujoco_rollouts.npz from the REAL reference engine (`mujoco.mj_step`), wherever it is importable.

This image has no `mujoco` (pinned ==3.11.0 in the reference's requirements.txt and absent from /opt/wheelhouse), so the
script cannot run here: `python tools/dump_mujoco_goldens.py --check` says so and exits 0. On a machine with
`pip install mujoco==3.11.0` and a dm_control checkout (default /root/reference, or --reference PATH):

    python tools/dump_mujoco_goldens.py            # writes tests/golden/mujoco_rollouts.npz
    python -m pytest tests/test_golden_rollouts.py # now also checks oracle AND CUDA path against real mj_step

It uses the SAME cases, seeds, start states and action tapes as tools/make_golden_rollouts.py (`inputs()` is shared),
loads the same reference XML the model fixtures were compiled from (tools/make_model_fixtures.py builds the exact XML
strings), and steps with the reference's legacy ordering (dm_control/mujoco/engine.py:147-162:
mj_step2, mj_step x (n-1), mj_step1). Per case it stores what the oracle file stores: final qpos / qvel / sensordata,
the ncon trace per control step and the final (geom1, geom2) list per environment.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
OUT = os.path.join(ROOT, 'tests', 'golden', 'mujoco_rollouts.npz')


def find_mujoco():
  for extra in (None, os.path.join(ROOT, 'baseline', '_ref')):
    if extra and os.path.isdir(extra) and extra not in sys.path:
      sys.path.insert(0, extra)
    try:
      import mujoco
      return mujoco
    except Exception:
      continue
  return None


def reference_xml(name, reference):
  """(xml string or bytes, asset dir) of a golden case, built exactly as the model fixtures were."""
  import make_model_fixtures as mf
  from dm_control_b200 import testing_models as tm
  base = name[:-len('_floor')] if name.endswith('_floor') else name
  suite_dir = os.path.join(reference, 'dm_control', 'suite')
  mf.REF = suite_dir
  if base in tm.XML:
    return tm.XML[base], None
  if base == 'quadruped':
    return mf.quadruped_walk_xml(), suite_dir
  if base == 'cmu_humanoid':
    return mf.cmu_humanoid_flat_xml(), None
  return open(os.path.join(suite_dir, base + '.xml'), 'rb').read(), suite_dir


def load_model(mujoco, name, reference):
  xml, base_dir = reference_xml(name, reference)
  assets = {}
  if base_dir:        # the suite XMLs include ./common/*.xml
    common = os.path.join(base_dir, 'common')
    for f in os.listdir(common):
      assets['./common/' + f] = open(os.path.join(common, f), 'rb').read()
  if isinstance(xml, bytes):
    xml = xml.decode()
  return mujoco.MjModel.from_xml_string(xml, assets or None)


def rollout(mujoco, name, nsub, nsteps, reference):
  import make_golden_rollouts as mg
  ours, q0, v0, tape = mg.inputs(name, nsteps)
  m = load_model(mujoco, name, reference)
  assert (m.nq, m.nv, m.nu, m.ngeom) == (ours.nq, ours.nv, ours.nu, ours.ngeom), (name, m.nq, m.nv, m.nu, m.ngeom)
  qs, vs, sens, pairs = [], [], [], []
  trace = np.zeros((nsteps, mg.B), np.int32)
  for e in range(mg.B):
    d = mujoco.MjData(m)
    d.qpos[:] = q0[e]; d.qvel[:] = v0[e]
    mujoco.mj_forward(m, d)
    for t in range(nsteps):
      d.ctrl[:] = tape[t, e]
      # legacy ordering of one control step (engine.py:147-162)
      if m.opt.integrator == mujoco.mjtIntegrator.mjINT_RK4:
        for _ in range(nsub):
          mujoco.mj_step(m, d)
      else:
        mujoco.mj_step2(m, d)
        for _ in range(nsub - 1):
          mujoco.mj_step(m, d)
      mujoco.mj_step1(m, d)
      trace[t, e] = d.ncon
    qs.append(d.qpos.copy()); vs.append(d.qvel.copy()); sens.append(d.sensordata.copy())
    pairs.append(np.array([[c.geom1, c.geom2] for c in d.contact[:d.ncon]], dtype=np.int32).reshape(-1, 2))
  return dict(qpos=np.stack(qs), qvel=np.stack(vs), sensordata=np.stack(sens), ncon=trace[-1].copy(), ncon_trace=trace,
              pairs=mg.pack_pairs(pairs))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--reference', default='/root/reference')
  ap.add_argument('--check', action='store_true', help='only report whether MuJoCo is importable')
  args = ap.parse_args()
  mujoco = find_mujoco()
  if mujoco is None:
    print('mujoco is not importable here: tests/golden/mujoco_rollouts.npz not written (parity stays oracle-only)')
    return 0 if args.check else 1
  print('mujoco', mujoco.__version__)
  if args.check:
    return 0
  import make_golden_rollouts as mg
  out = {'mujoco_version': np.array(mujoco.__version__)}
  for name, nsub, nsteps in mg.CASES:
    r = rollout(mujoco, name, nsub, nsteps, args.reference)
    for k, v in r.items():
      out[f'{name}_{k}'] = v
    print(name, 'final ncon', r['ncon'].tolist())
  np.savez_compressed(OUT, **out)
  print('wrote', OUT)
  return 0


if __name__ == '__main__':
  sys.exit(main())
