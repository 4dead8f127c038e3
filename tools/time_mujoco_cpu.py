"""CPU arm through the REAL reference (dm_control + mujoco), for machines that have them.

This image has neither `mujoco` nor dm_control's other dependencies, so `bench.py --impl reference` times the restated
oracle (kind "port"). Where `import mujoco` and `from dm_control import suite` work, bench.py calls `time_reference`
below instead and reports kind "reference": the unmodified `suite.load('humanoid', 'run')` environment stepped through
its own public `Environment.step` (rl/control.py:99-127) — one process per core, the same seeded start states,
settle and timed window as the GPU arm. It cannot be exercised here; it is kept deliberately small.
"""
from __future__ import annotations

import os
import time


def available():
  try:
    import mujoco  # noqa: F401
    from dm_control import suite  # noqa: F401
    return True
  except Exception:
    return False


def _worker(args):
  t, nenv, warmup_steps, timed_steps, start_at, root = args
  import sys
  sys.path.insert(0, root)
  import numpy as np
  from dm_control import suite
  from dm_control_b200 import testing_models as tm
  model = tm.load('humanoid')
  q0, v0 = tm.initial_states(model, 'humanoid', nenv, 7000 + t)
  envs = []
  for e in range(nenv):
    env = suite.load('humanoid', 'run')
    env.reset()
    with env.physics.reset_context():
      env.physics.data.qpos[:] = q0[e]
      env.physics.data.qvel[:] = v0[e]
    envs.append(env)
  nu = envs[0].action_spec().shape[0]
  tape = np.random.RandomState(100 + t).uniform(-1, 1, (warmup_steps + timed_steps, nenv, nu))
  for k in range(warmup_steps):
    for j, env in enumerate(envs):
      env.step(tape[k, j])
  while time.time() < start_at:
    time.sleep(0.001)
  t0 = time.time()
  for k in range(warmup_steps, warmup_steps + timed_steps):
    for j, env in enumerate(envs):
      env.step(tape[k, j])
  return t0, time.time()


def time_reference(warmup_steps, timed_steps, nenv_per_proc, procs=None):
  """-> (env_steps_per_s, cores, sample description, ms per env-step)."""
  import multiprocessing as mp
  procs = procs or os.cpu_count() or 1
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  start_at = time.time() + 10.0 + 0.05 * procs
  with mp.get_context('fork').Pool(procs) as pool:
    spans = pool.map(_worker, [(t, nenv_per_proc, warmup_steps, timed_steps, start_at, root) for t in range(procs)], chunksize=1)
  dt = max(b for _, b in spans) - min(a for a, _ in spans)
  nenv = procs * nenv_per_proc
  sample = (f'{nenv} envs ({nenv_per_proc}/process x {procs} processes) x {timed_steps} env-steps after {warmup_steps} '
            f'warm-up through dm_control suite.load(humanoid, run).step, seeded states, uniform(-1,1) actions')
  return nenv * timed_steps / dt, procs, sample, dt * 1e3 / timed_steps
