"""Regression goldens: seeded rollouts of the CPU oracle (NOT MuJoCo outputs — MuJoCo is absent from this image).

They freeze today's oracle behaviour so that an accidental change to oracle/mjoracle.cpp or to the MJCF compiler shows
up as a test failure on the CPU suite (tests/test_golden_rollouts.py) instead of silently moving the parity target.
When a real MuJoCo 3.11 is available, regenerate this file from `mujoco.mj_step` with the same seeds and tapes
(`--mujoco`, not implemented here because it cannot be exercised in this container).

Run:  python tools/make_golden_rollouts.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dm_control_b200 import testing_models as tm
from oracle import oracle as om

CASES = (('cartpole', 1, 40), ('cheetah', 1, 60), ('humanoid', 5, 16), ('quadruped', 4, 12), ('pendulum_free', 2, 30), ('cmu_humanoid', 6, 5))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'oracle_rollouts.npz')


def rollout(name, nsub, nsteps, B=3, seed=21):
  model = tm.load(name)
  q0, v0 = tm.initial_states(model, name, B, seed)
  tape = np.random.RandomState(seed + 1).uniform(-1, 1, (nsteps, B, model.nu))
  qs, vs, ncons, pairs = [], [], [], []
  for e in range(B):
    o = om.OraclePhysics(model)
    o.qpos[:] = q0[e]; o.qvel[:] = v0[e]; o.forward()
    for t in range(nsteps):
      o.ctrl[:] = tape[t, e]; o.control_step(nsub)
    qs.append(o.qpos.copy()); vs.append(o.qvel.copy()); ncons.append(o.ncon)
    pairs.append(np.array([[c.geom1, c.geom2] for c in o.contact], dtype=np.int32).reshape(-1, 2))
  return np.stack(qs), np.stack(vs), np.array(ncons, np.int32), pairs


def main():
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  out = {}
  for name, nsub, nsteps in CASES:
    q, v, n, pairs = rollout(name, nsub, nsteps)
    out[f'{name}_qpos'], out[f'{name}_qvel'], out[f'{name}_ncon'] = q, v, n
    out[f'{name}_pairs'] = np.concatenate(pairs) if sum(len(p) for p in pairs) else np.zeros((0, 2), np.int32)
    print(name, 'ncon', n.tolist(), 'max|q|', float(np.abs(q).max()))
  np.savez_compressed(OUT, **out)
  print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
  main()
