"""-m gpu: the batched suite tasks behave like the reference's conformance tests demand (suite/suite_test.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
from conftest import DEV   # noqa: E402  ('cuda'; 'cpu' only under B200MJ_EMULATE_GPU=1)


@pytest.mark.parametrize('domain,task,nu,obs_dim', [('cartpole', 'swingup', 1, 5), ('cheetah', 'run', 6, 17),
                                                     ('humanoid', 'run', 21, 67), ('quadruped', 'walk', 12, 78)])
def test_task_conformance(domain, task, nu, obs_dim):
  from dm_control_b200 import suite
  B = 64
  env = suite.load(domain, task, batch=B, seed=3)
  ts = env.reset()
  assert ts.reward is None and int(ts.step_type[0]) == 0
  g = torch.Generator(device=DEV).manual_seed(0)
  dims, seen = None, []
  for t in range(12):
    a = torch.rand(B, nu, generator=g, device=DEV, dtype=torch.float64) * 2 - 1     # suite_test.py:32-45 policy
    ts = env.step(a)
    flat = torch.cat([v.reshape(B, -1) for v in ts.observation.values()], dim=1)
    assert flat.shape == (B, obs_dim), flat.shape                                       # SURVEY §8a observation sizes
    assert bool(torch.isfinite(flat).all())                                              # suite_test.py:148-167
    assert float(ts.reward.min()) >= 0.0 and float(ts.reward.max()) <= 1.0
    seen.append(flat)
  assert float((seen[-1] - seen[0]).abs().max()) > 1e-6                                  # observations are not constant
  assert float(seen[0].std(dim=0).max()) > 1e-6                                          # initial states are randomised across envs


def test_same_seed_same_trajectory():
  # suite/suite_test.py:169-185
  from dm_control_b200 import suite
  outs = []
  for _ in range(2):
    env = suite.load('cheetah', 'run', batch=16, seed=11)
    env.reset()
    g = torch.Generator(device=DEV).manual_seed(5)
    for _ in range(5):
      ts = env.step(torch.rand(16, 6, generator=g, device=DEV, dtype=torch.float64) * 2 - 1)
    outs.append(torch.cat([v.reshape(16, -1) for v in ts.observation.values()], dim=1))
  assert torch.equal(outs[0], outs[1])


def test_humanoid_reward_matches_numpy_formula():
  """The torch reward equals the reference formula (suite/humanoid.py:183-207) evaluated in numpy on the same fields."""
  from dm_control_b200 import suite
  env = suite.load('humanoid', 'run', batch=32, seed=1)
  env.reset()
  g = torch.Generator(device=DEV).manual_seed(2)
  for _ in range(8):
    ts = env.step(torch.rand(32, 21, generator=g, device=DEV, dtype=torch.float64) * 2 - 1)
  p = env.physics
  head = p.head_height().cpu().numpy(); up = p.torso_upright().cpu().numpy(); ctrl = p.control().cpu().numpy()
  com = p.center_of_mass_velocity().cpu().numpy()
  def tol(x, lo, hi, margin, sig, vam):
    x = np.asarray(x); inb = (lo <= x) & (x <= hi); d = np.where(x < lo, lo - x, x - hi) / margin
    if sig == 'gaussian': s = np.exp(-0.5 * (d * np.sqrt(-2 * np.log(vam))) ** 2)
    elif sig == 'linear': sx = d * (1 - vam); s = np.where(abs(sx) < 1, 1 - sx, 0.0)
    else: sx = d * np.sqrt(1 - vam); s = np.where(abs(sx) < 1, 1 - sx ** 2, 0.0)
    return np.where(inb, 1.0, s)
  standing = tol(head, 1.4, np.inf, 0.35, 'gaussian', 0.1)
  upright = tol(up, 0.9, np.inf, 1.9, 'linear', 0)
  small = (4 + tol(ctrl, 0, 0, 1, 'quadratic', 0).mean(1)) / 5
  move = (5 * tol(np.linalg.norm(com[:, :2], axis=1), 10, np.inf, 10, 'linear', 0) + 1) / 6
  np.testing.assert_allclose(ts.reward.cpu().numpy(), small * standing * upright * move, rtol=1e-12, atol=1e-14)


def test_time_limit_and_auto_reset():
  """rl/control.py:99-127 semantics, batched: the step that reaches the limit is LAST; the next call re-initialises."""
  from dm_control_b200 import suite, control
  env = suite.load('cartpole', 'swingup', batch=8, seed=0, time_limit=0.05)   # 5 control steps of 0.01 s
  env.reset()
  a = torch.zeros(8, 1, dtype=torch.float64, device=DEV)
  types = []
  for _ in range(7):
    ts = env.step(a)
    types.append(int(ts.step_type[0]))
  assert types == [control.MID] * 4 + [control.LAST] + [control.MID] * 2
  assert float(env.physics.data.time.max()) == pytest.approx(0.02)          # two steps into the second episode
  # masked reset keeps the other environments untouched
  env2 = suite.load('cheetah', 'run', batch=6, seed=2)
  env2.reset()
  before = env2.physics.get_state().clone()
  mask = torch.tensor([True, False, False, True, False, False], device=DEV)
  env2.task.initialize_episode(env2.physics, mask)
  after = env2.physics.get_state()
  assert torch.equal(after[~mask], before[~mask])
  assert not torch.equal(after[mask], before[mask])


@pytest.mark.skipif(DEV != 'cuda', reason='CUDA graphs: device only')
@pytest.mark.parametrize('domain,task,nu', [('humanoid', 'run', 21), ('cheetah', 'run', 6)])
def test_whole_step_graph_equals_eager(domain, task, nu):
  """BatchedEnvironment(graph_step=True): the control step replayed as one CUDA graph (physics launches on the engine's
  streams + task ops) gives the same trajectory, bit for bit, as the eager path — across the capture itself, the
  reuse / no-reuse flag combinations and a reset in the middle."""
  from dm_control_b200 import suite
  outs = []
  for graphed in (False, True):
    env = suite.load(domain, task, batch=32, seed=4)
    env.physics.check_errors = False
    env._graph_step = graphed
    env.reset()
    g = torch.Generator(device=DEV).manual_seed(9)
    rows = []
    for t in range(14):
      if t == 8:
        env.reset()
      a = torch.rand(32, nu, generator=g, device=DEV, dtype=torch.float64) * 2 - 1
      ts = env.step(a)
      rows.append(torch.cat([v.reshape(32, -1) for v in ts.observation.values()] + [ts.reward[:, None]], dim=1).clone())
    outs.append((torch.stack(rows), env.physics.get_state().clone()))
    if graphed:
      assert any(isinstance(v, tuple) for v in env._step_graphs.values())      # a graph was really captured and replayed
  assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_bind_prefix_free_access_and_lazy_forward():
  """`physics.bind(kind, names)` (mjcf/physics.py:516-652): prefix-free model/data access per element, ragged joint /
  sensor fields, and the dirty contract — a qpos write through a binding makes the next derived read run forward()."""
  from dm_control_b200 import testing_models as tm
  from dm_control_b200.physics import BatchedPhysics
  phys = BatchedPhysics(tm.load('humanoid').copy(), batch=3)      # a private copy: the test edits the model
  m = phys.model
  head, hands = phys.bind('body', 'head'), phys.bind('body', ['left_hand', 'right_hand'])
  assert head.element_id == m.name2id('head', 'body') and tuple(hands.xpos.shape) == (3, 2, 3) and tuple(head.xpos.shape) == (3, 3)
  assert torch.equal(head.xpos, phys.named.data.xpos['head'])
  assert abs(float(head.mass) - float(np.asarray(m.body_mass)[head.element_id])) == 0.0
  knee = phys.bind('joint', 'right_knee')
  assert tuple(knee.range.shape) == (2,) and tuple(knee.qpos.shape) == (3, 1)
  root = phys.bind('joint', 'root')
  assert tuple(root.qpos.shape) == (3, 7) and tuple(root.qvel.shape) == (3, 6)
  z0 = head.xpos[:, 2].clone()
  assert not phys.is_dirty
  q = root.qpos.clone(); q[:, 2] += 0.25
  root.qpos = q                                   # state write through the binding -> dirty
  assert phys.is_dirty
  z1 = head.xpos[:, 2]                            # derived read -> forward() runs lazily, once
  assert not phys.is_dirty and float((z1 - z0 - 0.25).abs().max()) < 1e-12
  sens = phys.bind('sensor', 'torso_subtreelinvel')
  assert tuple(sens.sensordata.shape) == (3, 3)
  act = phys.bind('actuator', ['abdomen_y', 'abdomen_z'])
  act.ctrl = torch.tensor([0.5, -0.5], dtype=torch.float64)
  assert phys.data.ctrl[:, m.name2id('abdomen_y', 'actuator')].tolist() == [0.5] * 3 and not phys.is_dirty
  floor = phys.bind('geom', 'floor')
  v = m._version
  floor.friction = [0.9, 0.005, 0.0001]           # model write: tracked (re-uploaded before the next call)
  assert m._version == v + 1 and float(np.asarray(m.geom_friction).reshape(-1, 3)[floor.element_id, 0]) == 0.9
  phys.step()


def test_physics_pickles_with_model_and_state():
  """engine_test.py:574-589 (`pickle.loads(pickle.dumps(physics))`): model and state travel, the copy steps in lock-step."""
  import pickle
  from dm_control_b200 import testing_models as tm
  from dm_control_b200.physics import BatchedPhysics
  a = BatchedPhysics(tm.load('cheetah'), batch=5, outputs=('sensordata', 'xpos'))
  q0, v0 = tm.initial_states(a.model, 'cheetah', 5, 4)
  a.data.qpos.copy_(torch.as_tensor(q0)); a.data.qvel.copy_(torch.as_tensor(v0)); a.forward()
  a.set_control(torch.full((6,), 0.3, dtype=torch.float64)); a.step(7)
  b = pickle.loads(pickle.dumps(a))
  assert b.batch == 5 and torch.equal(a.get_state(), b.get_state()) and torch.equal(a.data.qacc_warmstart, b.data.qacc_warmstart)
  assert torch.equal(a.data.xpos, b.data.xpos) and not hasattr(b.data, 'xmat')
  for p in (a, b):
    p.step(5)
  assert torch.equal(a.get_state(), b.get_state()) and torch.equal(a.data.sensordata, b.data.sensordata)
