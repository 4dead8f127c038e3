"""Regression goldens: seeded rollouts of the CPU oracle (NOT MuJoCo outputs — MuJoCo is absent from this image).

They freeze today's oracle behaviour so that an accidental change to oracle/mjoracle.cpp or to the MJCF compiler shows
up as a test failure on the CPU suite (tests/test_golden_rollouts.py) instead of silently moving the parity target.
`tools/dump_mujoco_goldens.py` writes the same cases (same seeds, start states and action tapes) from the real
`mujoco.mj_step` wherever MuJoCo is importable; the tests prefer that file when it is present.

Every case ends IN CONTACT (quadruped and CMU humanoid are rolled out until they lie / stand on the floor), and the
file carries, per case: final qpos / qvel / sensordata, the per-control-step ncon trace, and the final contact list
(geom1, geom2) of every environment — what "contact-pair indexing bit-exact" is checked against.

Run:  python tools/make_golden_rollouts.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dm_control_b200 import testing_models as tm
from oracle import oracle as om

# (model, physics steps per control step, control steps)
CASES = (('cartpole', 1, 40), ('cheetah', 1, 60), ('humanoid', 5, 16), ('quadruped', 4, 12), ('quadruped_floor', 4, 30), ('pendulum_free', 2, 30), ('convex_zoo_floor', 5, 18),
         ('cmu_humanoid', 6, 24))
B, SEED = 3, 21
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'oracle_rollouts.npz')


def inputs(name, nsteps):
  """Start states and action tape of a case: the single definition both golden writers and the tests use."""
  model = tm.load(name)
  q0, v0 = tm.initial_states(model, name, B, SEED)
  tape = np.random.RandomState(SEED + 1).uniform(-1, 1, (nsteps, B, model.nu))
  return model, q0, v0, tape


def pack_pairs(pairs):
  """[(n_e, 2)] per environment -> one (sum n_e, 2) array; split again with the final ncon."""
  return np.concatenate(pairs).astype(np.int32) if sum(len(p) for p in pairs) else np.zeros((0, 2), np.int32)


def rollout(name, nsub, nsteps):
  model, q0, v0, tape = inputs(name, nsteps)
  qs, vs, sens, pairs = [], [], [], []
  trace = np.zeros((nsteps, B), np.int32)
  qtrace = np.zeros((nsteps, B, model.nq))
  # first control step after which a contact of a pair WITHOUT closed form (MPR: an ellipsoid, a non-plane cylinder or
  # two boxes involved) is active, per environment (nsteps = never). Those contacts are a discontinuous function of the
  # pose (include/b200mj_convex.h), so beyond that step implementations that differ in rounding — fused multiply-add on
  # the GPU — may legitimately part ways; the strict comparisons stop there (tests/test_golden_rollouts.py).
  gt = np.asarray(model.geom_type)
  mpr_step = np.full(B, nsteps, np.int32)
  for e in range(B):
    o = om.OraclePhysics(model)
    o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward()
    for t in range(nsteps):
      o.ctrl[:] = tape[t, e]; o.control_step(nsub)
      trace[t, e] = o.ncon
      qtrace[t, e] = o.qpos
      if mpr_step[e] == nsteps and any(gt[c.geom1] != 0 and (gt[c.geom1] in (4, 5) or gt[c.geom2] in (4, 5) or (gt[c.geom1] == 6 and gt[c.geom2] == 6))
                                       for c in o.contact):
        mpr_step[e] = t
    qs.append(o.qpos.copy()); vs.append(o.qvel.copy()); sens.append(np.array(o.sensordata, dtype=np.float64).copy())
    pairs.append(np.array([[c.geom1, c.geom2] for c in o.contact], dtype=np.int32).reshape(-1, 2))
  return dict(qpos=np.stack(qs), qvel=np.stack(vs), sensordata=np.stack(sens), ncon=trace[-1].copy(), ncon_trace=trace,
              pairs=pack_pairs(pairs), qpos_trace=qtrace, mpr_step=mpr_step)


def main():
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  out = {}
  for name, nsub, nsteps in CASES:
    r = rollout(name, nsub, nsteps)
    for k, v in r.items():
      out[f'{name}_{k}'] = v
    print(name, 'final ncon', r['ncon'].tolist(), 'pairs', len(r['pairs']), 'max|q|', float(np.abs(r['qpos']).max()))
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
