"""-m gpu: BASELINE.json config 5 — the batched composer CMU-humanoid run-through-corridor task.

Pins restated from the reference's own tests (locomotion/tasks/corridors_test.py:40-130) and parity of the
per-environment corridors (wall boxes as per-environment geoms) against the CPU oracle on per-environment MODELS, which
is what the reference builds when it recompiles every episode (composer/environment.py:378-383)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
from conftest import DEV   # noqa: E402


def _env(B, seed=1):
  from dm_control_b200 import locomotion
  return locomotion.load('cmu_humanoid_run_walls', batch=B, seed=seed)


def test_walker_is_reinitialised_upright_at_the_spawn_position():
  # corridors_test.py:40-99 (no rotation): joints at qpos0, root at upright pose + spawn offset (0.5, 0, 0)
  env = _env(3)
  env.reset()
  g = torch.Generator(device=DEV).manual_seed(0)
  for _ in range(3):
    env.step(torch.rand(3, 56, generator=g, device=DEV, dtype=torch.float64) * 2 - 1)
  env.reset()
  q = env.physics.data.qpos.cpu().numpy()
  assert np.array_equal(q[:, 7:], np.zeros_like(q[:, 7:]))
  np.testing.assert_array_equal(q[:, :3], np.tile([0.5, 0.0, 0.94], (3, 1)))
  up = np.array([0.859, 1.0, 1.0, 0.859]); up /= np.linalg.norm(up)
  np.testing.assert_allclose(q[:, 3:7], np.tile(up, (3, 1)), atol=1e-15)
  assert float(env.physics.data.qvel.abs().max()) == 0.0


def test_termination_and_discount():
  # corridors_test.py:101-130: upright for the first steps -> no termination, discount 1; an inverted walker dropped
  # onto the ground registers a non-foot ground contact -> termination, discount 0
  env = _env(2)
  env.reset()
  zero = torch.zeros(2, 56, dtype=torch.float64, device=DEV)
  for _ in range(5):
    ts = env.step(zero)
    assert not bool(env.task.should_terminate_episode(env.physics).any())
    assert ts.discount.tolist() == [1.0, 1.0] and ts.step_type.tolist() == [1, 1]
  phys = env.physics
  phys.data.qpos[0, 2] = 1.2                     # env 0 only: upside down, just above the ground
  phys.data.qpos[0, 3:7] = torch.tensor([0.0, 1.0, 0.0, 0.0], dtype=torch.float64)
  phys.data.qvel[0].zero_()
  phys.forward()
  for _ in range(400):
    if int(phys.data.ncon[0]) > 0:
      break
    phys.step()
  assert int(phys.data.ncon[0]) > 0
  ts = env.step(zero)
  assert env.task.should_terminate_episode(phys).tolist() == [True, False]
  assert ts.discount.tolist() == [0.0, 1.0] and ts.step_type.tolist() == [2, 1]
  ts = env.step(zero)                             # env 0 is reset in place (upright again, its clock restarted), env 1 carries on
  assert int(ts.step_type[0]) == 1
  assert 0.8 < float(phys.data.qpos[0, 2]) < 1.0 and abs(float(phys.data.time[0]) - 0.03) < 1e-12 and float(phys.data.time[1]) > 0.2


def test_walls_are_redrawn_per_environment_and_per_episode():
  # arenas/corridors.py:394-440 with basic_cmu_2019.py:40-48: 25 walls at x = 2 + 4k, width U(1, 7), sides alternate
  env = _env(4, seed=7)
  env.reset()
  p, s = env.physics.data.var_geom_pos.cpu().numpy(), env.physics.data.var_geom_size.cpu().numpy()
  assert p.shape == (4, 25, 3)
  np.testing.assert_array_equal(p[:, :, 0], np.tile(2.0 + 4.0 * np.arange(25), (4, 1)))
  w = 2 * s[:, :, 1]
  assert w.min() >= 1.0 and w.max() <= 7.0 and np.std(w) > 0.5 and not np.allclose(w[0], w[1])
  np.testing.assert_allclose(p[:, :, 1], np.where(np.arange(25) % 2 == 0, 1, -1) * (10.0 - w) / 2)
  np.testing.assert_array_equal(s[:, :, 0], 0.08); np.testing.assert_array_equal(p[:, :, 2], 1.5)
  first = p.copy()
  env._reset_envs(torch.tensor([True, False, False, False], device=DEV))
  p2 = env.physics.data.var_geom_pos.cpu().numpy()
  assert not np.allclose(p2[0], first[0]) and np.array_equal(p2[1:], first[1:])


def test_corridor_rollout_matches_per_environment_oracle_models(oracle_mod):
  """Every environment against an oracle whose MODEL carries that environment's walls: state, contact pairs,
  subtree_linvel (the reward's input) and the termination flag of the reference's contact rule."""
  from oracle import oracle as om
  B, nstep = 3, 10
  env = _env(B, seed=3)
  env.reset()
  phys, task = env.physics, env.task
  # walk the walkers into their first wall: start 10 cm in front of it, leaning on it
  pos, size = phys.data.var_geom_pos.cpu().numpy(), phys.data.var_geom_size.cpu().numpy()
  q = phys.data.qpos.clone()
  q[:, 0] = 2.0 - 0.3
  q[:, 1] = torch.as_tensor(pos[:, 0, 1], device=DEV)
  phys.data.qpos.copy_(q); phys.forward()
  tape = np.random.RandomState(5).uniform(-1, 1, (nstep, B, 56))
  q0 = phys.data.qpos.cpu().numpy().copy()
  fails = []
  for t in range(nstep):
    env._reset_next.zero_()                      # keep stepping through terminations: this test compares physics
    env.step(torch.as_tensor(tape[t], device=DEV))
    fails.append(task.should_terminate_episode(phys).cpu().numpy().copy())
  m = phys.model
  nonfoot = task.walker.nonfoot_geom.cpu().numpy(); ground = task.walker.ground_geom
  touched_wall = 0
  for e in range(B):
    me = m.copy()
    for k, gid in enumerate(task.wall_geoms):
      me.fields['geom_pos'].reshape(-1, 3)[gid] = pos[e, k]; me.fields['geom_size'].reshape(-1, 3)[gid] = size[e, k]
      me.fields['geom_rbound'].reshape(-1)[gid] = np.linalg.norm(size[e, k])
    me.touch()
    o = om.OraclePhysics(me)
    o.qpos[:] = q0[e]; o.forward()
    for t in range(nstep):
      o.ctrl[:] = tape[t, e]; o.control_step(env.n_sub_steps)
      bad = any((c.geom1 == ground and nonfoot[c.geom2]) or (c.geom2 == ground and nonfoot[c.geom1]) for c in o.contact)
      low = bool((np.asarray(o.xpos).reshape(-1, 3)[task.walker.end_effectors, 2] < -0.5).any())
      assert bool(fails[t][e]) == (bad or low), (e, t)
    assert np.abs(phys.data.qpos[e].cpu().numpy() - o.qpos).max() < 1e-6, e
    n = int(phys.data.ncon[e]); assert n == o.ncon
    pairs = [(int(a), int(b)) for a, b in phys.data.contact_geom[e].reshape(-1, 2)[:n].cpu().numpy()]
    assert pairs == [(c.geom1, c.geom2) for c in o.contact]
    touched_wall += any(g in task.wall_geoms for pr in pairs for g in pr)
    o.subtree_vel()                                # walker.after_substep: mj_subtreeVel (legacy_base.py:179-186)
    sl = np.asarray(o.subtree_linvel).reshape(-1, 3)[task.walker.root]
    assert np.abs(phys.data.subtree_linvel[e].reshape(-1, 3)[task.walker.root].cpu().numpy() - sl).max() < 1e-6
  assert int(phys.data.warning.sum()) == 0
