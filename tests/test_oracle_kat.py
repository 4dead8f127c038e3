"""CPU oracle vs the reference's own analytic known-answers (SURVEY.md §8c items 1,2,3,5,9).

These are the only numerical pins the reference tree holds for this path; trajectory parity against a real
MuJoCo 3.11.0 remains UNPINNED. Everything here runs without a GPU.
"""
import numpy as np
import pytest

from dm_control_b200 import mjcf_compile as mc
from dm_control_b200 import testing_models as tm

DSBL_CONTACT, DSBL_GRAVITY = 1 << 4, 1 << 6


def _oracle(oracle_mod, xml=None, name=None):
  model = mc.compile_xml(xml) if xml is not None else tm.load(name)
  return model, oracle_mod.OraclePhysics(model)


def test_readme_rest_depth(oracle_mod):
  # dm_control/mujoco/README.md:25-50: box+sphere body on a z-slide; after 1 s geom z = [0.19996362, 0.39996362]
  model, p = _oracle(oracle_mod, name='slide_box')
  p.qpos[0] = 0.5
  p.forward()
  np.testing.assert_allclose(p.geom_xpos[1], [0, 0, 0.8], atol=1e-12)
  np.testing.assert_allclose(p.geom_xpos[2], [0.2, 0.2, 1.0], atol=1e-12)
  while p.time < 1.0:
    p.step()
  assert abs(p.geom_xpos[1, 2] - 0.19996362) < 5e-9
  assert abs(p.geom_xpos[2, 2] - 0.39996362) < 5e-9


def test_contact_force_equals_weight(oracle_mod):
  # dm_control/mujoco/wrapper/core_test.py:393-416
  model, p = _oracle(oracle_mod, name='free_box')
  p.forward()
  p.step(500)
  normal = sum(p.contact_force(i)[0] for i in range(p.ncon))
  weight = 9.81 * model.body_mass[1]
  assert abs(model.body_mass[1] - 8.0) < 1e-12       # (0.2 m)^3 * 1000 kg/m^3
  assert p.ncon == 4
  assert abs(normal - weight) < 5e-8                  # assertAlmostEqual default: 7 places


SLIDING_CUBE = """
<mujoco>
  <option gravity="0 0 -9.81"/>
  <worldbody>
    <geom name="floor" type="plane" pos="0 0 0" size="10 10 0.1"/>
    <body name="cube" pos="0 0 0.1">
      <geom type="box" size="0.1 0.1 0.1" mass="1"/>
      <site name="cube_site" type="box" size="0.1 0.1 0.1"/>
      <joint type="slide"/>
    </body>
  </worldbody>
  <sensor><touch name="touch_sensor" site="cube_site"/></sensor>
</mujoco>"""


def test_disable_flags_touch_sensor(oracle_mod):
  # dm_control/mujoco/wrapper/core_test.py:291-330
  model, p = _oracle(oracle_mod, xml=SLIDING_CUBE)
  p.forward()
  p.step(100)
  assert abs(p.qvel[0]) < 5e-5
  assert abs(p.sensordata[0] - 9.81) < 5e-3
  flags = p.disableflags
  p.disableflags = flags | DSBL_CONTACT | DSBL_GRAVITY
  p.step(1)
  assert abs(p.qvel[0]) < 5e-5
  assert p.sensordata[0] == 0
  p.disableflags = flags | DSBL_CONTACT
  p.step(10)
  assert p.qvel[0] < -0.1


CARTPOLE_TIP = """
<mujoco><worldbody>
  <body name='cart'><joint type='slide' axis='1 0 0'/><geom name='cart' type='box' size='0.2 0.2 0.2'/>
    <body name='pole'><joint name='hinge' type='hinge' axis='0 1 0'/><geom name='mass' pos='0 0 .5' size='0.04'/></body>
  </body>
</worldbody></mujoco>"""


@pytest.mark.parametrize('qpos,expected_lin', [([0.0, 0.0], [1.5, 0, 0]), ([0.0, np.pi], [0.5, 0, 0])])
def test_object_velocity(oracle_mod, qpos, expected_lin):
  # dm_control/mujoco/wrapper/core_test.py:340-391 (world-frame cases): tip velocity from cvel
  model, p = _oracle(oracle_mod, xml=CARTPOLE_TIP)
  p.qpos[:] = qpos
  p.qvel[:] = [1.0, 1.0]
  p.step1()
  g = model.names['geom']['mass']
  b = model.geom_bodyid[g]
  cvel = p.cvel[b]
  r = p.geom_xpos[g] - p.subtree_com[model.body_rootid[b]]
  lin = cvel[3:] + np.cross(cvel[:3], r)
  np.testing.assert_allclose(lin, expected_lin, atol=5e-7)
  np.testing.assert_allclose(cvel[:3], [0, 1, 0], atol=5e-7)


ACCEL = """
<mujoco><worldbody>
  <geom type="plane" size="1 1 .1"/>
  <body pos="0 0 1"><joint type="slide" axis="0 0 1" range="-1 1" limited="true" stiffness="10000" damping="1000"/>
    <geom type="sphere" size=".1"/><site name="s"/></body>
</worldbody><sensor><accelerometer name="accelerometer" site="s"/></sensor></mujoco>"""


def test_accelerometer_reads_g_when_supported(oracle_mod):
  # dm_control/mujoco/engine_test.py:591-597: a body held still reads +g on its accelerometer z
  model, p = _oracle(oracle_mod, xml=ACCEL)
  p.forward()
  p.step(3000)   # settle on the stiff spring
  assert abs(p.qvel[0]) < 1e-6
  assert abs(p.sensordata[2] - 9.81) < 1e-4


def test_euler_update_closed_form(oracle_mod):
  # suite/lqr_solver.py:44-66 states the smooth Euler update in closed form; with damping treated implicitly
  # (MuJoCo's eulerdamp) one step of a 1-dof spring-damper is v' = v + h (m + h b)^-1 (-k q - b v + u), q' = q + h v'
  xml = """
  <mujoco><option timestep=".03"><flag constraint="disable" gravity="disable"/></option><worldbody>
    <body><joint name="j" type="slide" axis="0 1 0" stiffness="3" damping=".7"/><geom type="sphere" size=".1" mass="2"/></body>
  </worldbody><actuator><motor joint="j"/></actuator></mujoco>"""
  model, p = _oracle(oracle_mod, xml=xml)
  q, v, u, h, m_, k, b = 0.3, -0.2, 0.5, 0.03, 2.0, 3.0, 0.7
  p.qpos[0], p.qvel[0], p.ctrl[0] = q, v, u
  p.step(1)
  v1 = v + h * (-k * q - b * v + u) / (m_ + h * b)
  np.testing.assert_allclose(p.qvel[0], v1, rtol=1e-13)
  np.testing.assert_allclose(p.qpos[0], q + h * v1, rtol=1e-13)


def test_nstep_equals_repeated_step(oracle_mod):
  # dm_control/mujoco/engine_test.py:627-663 for the oracle itself (Euler and RK4)
  for name in ('cheetah', 'cartpole'):
    model = tm.load(name)
    q0, v0 = tm.initial_states(model, name, 1, 3)
    a, b = oracle_mod.OraclePhysics(model), oracle_mod.OraclePhysics(model)
    for o in (a, b):
      o.qpos[:] = q0[0]; o.qvel[:] = v0[0]; o.ctrl[:] = 0.3
    a.step(4)
    for _ in range(4):
      b.step(1)
    np.testing.assert_array_equal(a.qpos, b.qpos)
    np.testing.assert_array_equal(a.qvel, b.qvel)
