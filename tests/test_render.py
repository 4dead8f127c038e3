"""Rendering hand-off: the ray-cast `Physics.render` / `Camera` against the reference's own known-answers.

CPU tier: the numpy oracle (oracle/render_oracle.py) and the compiler's camera tables are held to
  * engine_test.py:64-80   testDepthRender        — nearest pixel 2.8 m, furthest 3.0 m (orthographic depth),
  * engine_test.py:82-131  testSegmentationRender — geom / site ids and mjtObj types in the four corners, -1 at the centre,
  * engine_test.py:228-275 testCameraMatrix       — a geom's centre, projected with `camera.matrix`, lands on a pixel that
                                                     the segmentation image labels with that geom (two model cameras and
                                                     the free camera, three image sizes),
  * engine_test.py:304-331 — the exceptions.
GPU tier (`-m gpu`): the CUDA kernel (b200mj_render through `BatchedPhysics.render`) reproduces the same pins and agrees
with the oracle pixel by pixel on the humanoid (trackcom camera) and on the CMU corridor's egocentric camera with
per-environment walls.

The scenes are the reference's XML strings; the engine needs at least one degree of freedom, so the GPU twins add an
invisible free body far away from the scene.
"""
import os

import numpy as np
import pytest

from dm_control_b200 import mjcf_compile as mc
from oracle import render_oracle as ro

PLANE_AND_BOX = """
<mujoco>
  <worldbody>
    <geom type="plane" pos="0 0 0" size="2 2 .1"/>
    <geom type="box" size=".1 .1 .1" pos="0 0 .1"/>
    <camera name="top" pos="0 0 3"/>
    %s
  </worldbody>
</mujoco>
"""
BOX_FOUR_CORNERS = """
<mujoco>
  <visual>
    <scale framelength="2"/>
  </visual>
  <worldbody>
    <geom name="box0" type="box" size=".2 .2 .2" pos="-1 1 .1"/>
    <geom name="box1" type="box" size=".2 .2 .2" pos="1 1 .1"/>
    <site name="box2" type="box" size=".2 .2 .2" pos="1 -1 .1"/>
    <site name="box3" type="box" size=".2 .2 .2" pos="-1 -1 .1"/>
    <camera name="top" pos="0 0 3"/>
    %s
  </worldbody>
</mujoco>
"""
TWO_GEOMS_TWO_CAMERAS = """
<mujoco>
  <visual>
    <global fovy="55"/>
  </visual>
  <worldbody>
    <light name="top" pos="0 0 1"/>
    <geom name="red" pos=".2 0 0" size=".005" rgba="1 0 0 1"/>
    <geom name="green" pos=".2 .2 .1" size=".005" rgba="0 1 0 1"/>
    <camera name="cam0" pos="1 .5 1" zaxis="1 .5 1" fovy="20"/>
    <camera name="cam1" pos=".1 .1 1" xyaxes="1 1 0 -1 0 0"/>
    %s
  </worldbody>
</mujoco>
"""
# The free camera's default pose hangs on mjModel.stat.{center, extent}, which MuJoCo derives at compile time and this
# compiler only estimates (geom bounding spheres at qpos0); the camera-matrix pin does not depend on the pose, so the
# free-camera case uses an explicit one that sees both geoms.
FREE_POSE = (np.array([0.2, 0.1, 0.05]), 0.6, 90.0, -45.0)
# the engine needs nv > 0: an invisible free body, out of every camera's sight, gravity compensated by nothing (it falls; irrelevant)
DUMMY = '<body name="dummy" pos="0 0 -50"><freejoint/><geom name="dummy" size=".01" rgba="0 0 0 0" contype="0" conaffinity="0"/></body>'


def _static_frames(model):
  """World frames of geoms / sites attached to the world body (all of the pinned scenes')."""
  q2m = lambda q: ro.quat_to_mat(q).reshape(-1)
  gx = np.asarray(model.geom_pos).reshape(-1, 3); gm = np.stack([q2m(q) for q in np.asarray(model.geom_quat).reshape(-1, 4)])
  if model.nsite:
    sx = np.asarray(model.site_pos).reshape(-1, 3); sm = np.stack([q2m(q) for q in np.asarray(model.site_quat).reshape(-1, 4)])
  else:
    sx, sm = np.zeros((0, 3)), np.zeros((0, 9))
  return gx, gm, sx, sm


def _oracle_images(model, camera_id, height, width, free_pose=None):
  vis = model.vis
  gx, gm, sx, sm = _static_frames(model)
  xpos = np.zeros((model.nbody, 3)); xmat = np.tile(np.eye(3).reshape(-1), (model.nbody, 1))
  if camera_id == -1:
    lookat, dist, az, el = free_pose or (vis['stat_center'], 1.5 * vis['stat_extent'][0], 90.0, -45.0)
    cp, cm = ro.free_camera_pose(lookat, dist, az, el)
    fovy = float(vis['global_fovy'][0])
  else:
    cp, cm = ro.camera_pose(vis, camera_id, xpos, xmat, xpos)
    fovy = float(vis['cam_fovy'][camera_id])
  rgb, depth, seg = ro.render(vis, np.asarray(model.geom_type), np.asarray(model.geom_size).reshape(-1, 3), gx, gm,
                              np.asarray(model.site_type), np.asarray(model.site_size).reshape(-1, 3), sx, sm, cp, cm, fovy, height, width)
  return rgb, depth, seg, (cp, cm, fovy)


# ---- CPU tier: oracle + compiler against the reference's known-answers ------------------------------------------------

def test_oracle_depth_render_pin():
  model = mc.compile_xml(PLANE_AND_BOX % '')
  _, depth, _, _ = _oracle_images(model, model.name2id('top', 'camera'), 200, 200)
  np.testing.assert_approx_equal(depth.min(), 2.8, 3)      # engine_test.py:77-78
  np.testing.assert_approx_equal(depth.max(), 3.0, 3)      # engine_test.py:79-80 (depth is orthographic)


def _check_segmentation_pin(pixels, model):
  np.testing.assert_equal(pixels[95:105, 95:105, 0], -1)
  np.testing.assert_equal(pixels[95:105, 95:105, 1], -1)
  np.testing.assert_equal(pixels[15:25, 0:10, 1], ro.OBJ_GEOM)
  np.testing.assert_equal(pixels[15:25, 190:200, 1], ro.OBJ_GEOM)
  np.testing.assert_equal(pixels[190:200, 190:200, 1], ro.OBJ_SITE)
  np.testing.assert_equal(pixels[190:200, 0:10, 1], ro.OBJ_SITE)
  np.testing.assert_equal(pixels[15:25, 0:10, 0], model.name2id('box0', 'geom'))
  np.testing.assert_equal(pixels[15:25, 190:200, 0], model.name2id('box1', 'geom'))
  np.testing.assert_equal(pixels[190:200, 190:200, 0], model.name2id('box2', 'site'))
  np.testing.assert_equal(pixels[190:200, 0:10, 0], model.name2id('box3', 'site'))


def test_oracle_segmentation_render_pin():
  model = mc.compile_xml(BOX_FOUR_CORNERS % '')
  _, _, seg, _ = _oracle_images(model, 0, 200, 200)
  _check_segmentation_pin(seg, model)                     # engine_test.py:104-126


@pytest.mark.parametrize('camera_id,height,width', [('cam0', 200, 300), (1, 300, 200), (-1, 400, 400)])
def test_oracle_camera_matrix_pin(camera_id, height, width):
  model = mc.compile_xml(TWO_GEOMS_TWO_CAMERAS % '')
  cid = model.name2id(camera_id, 'camera') if isinstance(camera_id, str) else camera_id
  _, _, seg, (cp, cm, fovy) = _oracle_images(model, cid, height, width, free_pose=FREE_POSE)
  image, focal, rotation, translation = ro.camera_matrix(cp, cm, fovy, height, width)
  cam = image @ focal @ rotation @ translation
  gx = np.asarray(model.geom_pos).reshape(-1, 3)
  for geom_id in (0, 1):
    xs, ys, s = cam.dot(np.array([*gx[geom_id], 1.0]))      # xyz2pixels, engine_test.py:243-246
    row, column = int(round(ys / s)), int(round(xs / s))
    assert tuple(seg[row, column]) == (geom_id, ro.OBJ_GEOM)


def test_compiler_camera_tables():
  model = mc.compile_xml(TWO_GEOMS_TWO_CAMERAS % '')
  vis = model.vis
  assert vis['cam_fovy'].tolist() == [20.0, 55.0] and float(vis['global_fovy'][0]) == 55.0
  z = ro.quat_to_mat(vis['cam_quat'][0])[:, 2]
  np.testing.assert_allclose(z, np.array([1, .5, 1]) / 1.5, atol=1e-12)          # zaxis="1 .5 1"
  R = ro.quat_to_mat(vis['cam_quat'][1])
  np.testing.assert_allclose(R[:, 0], np.array([1, 1, 0]) / np.sqrt(2), atol=1e-12)
  np.testing.assert_allclose(vis['geom_rgba'][0], [1, 0, 0, 1])
  with pytest.raises(ValueError):
    model.name2id('nope', 'camera')


def test_ray_primitives_against_closed_forms():
  o = np.array([[0.0, 0, 5.0], [0.3, 0, 5.0], [2.0, 0, 5.0]]); d = np.tile([0.0, 0, -1.0], (3, 1))
  t, n = ro.ray_primitive(2, np.array([1.0, 0, 0]), o, d, 1e-9)
  np.testing.assert_allclose(t[:2], [4.0, 5 - np.sqrt(1 - 0.09)]); assert np.isinf(t[2])
  t, _ = ro.ray_primitive(3, np.array([0.5, 1.0, 0]), o, d, 1e-9)           # capsule along z: cap top at 1.5
  np.testing.assert_allclose(t[0], 3.5); assert np.isinf(t[2])
  t, _ = ro.ray_primitive(5, np.array([0.5, 1.0, 0]), o, d, 1e-9)           # cylinder: flat cap at 1
  np.testing.assert_allclose(t[:2], [4.0, 4.0])
  t, n = ro.ray_primitive(6, np.array([0.5, 0.5, 0.25]), o, d, 1e-9)
  np.testing.assert_allclose(t[:2], [4.75, 4.75]); np.testing.assert_allclose(n[0], [0, 0, 1])
  t, _ = ro.ray_primitive(4, np.array([1.0, 1.0, 2.0]), o, d, 1e-9)         # ellipsoid, semi-axis 2 along z
  np.testing.assert_allclose(t[0], 3.0)
  t, _ = ro.ray_primitive(0, np.array([1.0, 1.0, 0.1]), o, d, 1e-9)         # finite plane |x| <= 1
  np.testing.assert_allclose(t[:2], [5.0, 5.0]); assert np.isinf(t[2])
  t, _ = ro.ray_primitive(0, np.array([0.0, 0.0, 0.1]), o, -d, 1e-9)        # from behind / pointing away: no hit
  assert np.isinf(t).all()


def test_camera_modes_of_the_oracle():
  """fixed / track / trackcom / targetbody poses from body frames (mjtCamLight semantics) on a two-body toy model."""
  xml = """
  <mujoco>
    <worldbody>
      <body name="a" pos="0 0 1">
        <freejoint/>
        <geom name="ga" size=".1"/>
        <camera name="fixed" pos="0 -2 0" xyaxes="1 0 0 0 0 1"/>
        <camera name="track" pos="0 -2 0" xyaxes="1 0 0 0 0 1" mode="track"/>
        <camera name="trackcom" pos="0 -2 0" xyaxes="1 0 0 0 0 1" mode="trackcom"/>
        <camera name="target" pos="0 -2 1" mode="targetbody" target="b"/>
        <body name="b" pos="1 0 0"><joint type="hinge" axis="0 0 1"/><geom name="gb" size=".1" mass="3"/></body>
      </body>
    </worldbody>
  </mujoco>"""
  model = mc.compile_xml(xml)
  vis = model.vis
  assert vis['cam_mode'].tolist() == [0, 1, 2, 3] and int(vis['cam_targetbodyid'][3]) == model.name2id('b', 'body')
  # body a moved to (2, 0, 1) and yawed by 90 degrees; b hangs 1 m along a's x axis -> world (2, 1, 1)
  Rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
  xpos = np.array([[0, 0, 0], [2.0, 0, 1], [2.0, 1, 1]]); xmat = np.stack([np.eye(3).reshape(-1), Rz.reshape(-1), Rz.reshape(-1)])
  com = np.array([[0, 0, 0], [2.0, 0.75, 1.0], [2.0, 1, 1]])
  p, R = ro.camera_pose(vis, 0, xpos, xmat, com)                 # fixed: rides the body frame
  np.testing.assert_allclose(p, [4.0, 0, 1], atol=1e-12); np.testing.assert_allclose(R[:, 2], Rz @ np.array([0, -1.0, 0]), atol=1e-12)
  p, R = ro.camera_pose(vis, 1, xpos, xmat, com)                 # track: world offset from the body origin and world orientation at qpos0
  np.testing.assert_allclose(p, [2.0, -2, 1], atol=1e-12); np.testing.assert_allclose(R[:, 2], [0, -1.0, 0], atol=1e-12)
  p, R = ro.camera_pose(vis, 2, xpos, xmat, com)                 # trackcom: the same offset, from the subtree's centre of mass
  np.testing.assert_allclose(p, com[1] + vis['cam_poscom0'][2], atol=1e-12)
  m_a = 1000 * 4 / 3 * np.pi * 0.1 ** 3                          # default density; b's geom has mass 3 at x = 1
  np.testing.assert_allclose(vis['cam_poscom0'][2], [-3 / (3 + m_a), -2.0, 0.0], atol=1e-12)
  p, R = ro.camera_pose(vis, 3, xpos, xmat, com)                 # targetbody: -z points at b, x horizontal
  to_b = xpos[2] - p
  np.testing.assert_allclose(-R[:, 2], to_b / np.linalg.norm(to_b), atol=1e-12); assert abs(R[2, 0]) < 1e-12


def test_host_camera_poses_and_matrices_match_the_oracle():
  """`render.Camera.pose / matrices / matrix` (torch, the product's host code) on CPU tensors against the numpy restatement,
  for every camera mode and the free camera — no kernel involved (the physics object is a stub holding body frames)."""
  import types
  import torch
  from dm_control_b200 import render
  xml = """
  <mujoco>
    <visual><global fovy="50"/></visual>
    <worldbody>
      <body name="a" pos="0 0 1">
        <freejoint/>
        <geom name="ga" size=".1"/>
        <camera name="fixed" pos="0 -2 0.3" xyaxes="1 0 0 0 0.2 1" fovy="30"/>
        <camera name="track" pos="0 -2 0" xyaxes="1 0 0 0 0 1" mode="track"/>
        <camera name="trackcom" pos="1 -2 0" xyaxes="1 0 0 0 0 1" mode="trackcom"/>
        <camera name="target" pos="0 -2 1" mode="targetbody" target="b"/>
        <camera name="targetcom" pos="0.5 -2 1" mode="targetbodycom" target="a"/>
        <body name="b" pos="1 0 0"><joint type="hinge" axis="0 0 1"/><geom name="gb" size=".1" mass="3"/></body>
      </body>
    </worldbody>
  </mujoco>"""
  model = mc.compile_xml(xml)
  B, rs = 3, np.random.RandomState(0)
  xpos = rs.uniform(-1, 1, (B, model.nbody, 3)); com = rs.uniform(-1, 1, (B, model.nbody, 3))
  xmat = np.zeros((B, model.nbody, 9))
  for e in range(B):
    for b in range(model.nbody):
      q = rs.randn(4); xmat[e, b] = ro.quat_to_mat(q / np.linalg.norm(q)).reshape(-1)
  t = lambda a: torch.as_tensor(a.reshape(B, -1).copy())
  data = types.SimpleNamespace(xpos=t(xpos), xmat=t(xmat), subtree_com=t(com), geom_xpos=torch.zeros(B, model.ngeom * 3, dtype=torch.float64),
                               geom_xmat=torch.zeros(B, model.ngeom * 9, dtype=torch.float64))
  phys = types.SimpleNamespace(model=model, data=data, batch=B, device=torch.device('cpu'), is_dirty=False)
  H, W = 48, 64
  for cid in range(5):
    cam = render.Camera(phys, H, W, cid)
    pos, mat = cam.pose()
    full = cam.matrix.numpy()
    for e in range(B):
      p_o, R_o = ro.camera_pose(model.vis, cid, xpos[e], xmat[e], com[e])
      np.testing.assert_allclose(pos[e].numpy(), p_o, atol=1e-12); np.testing.assert_allclose(mat[e].numpy(), R_o, atol=1e-12)
      image, focal, rotation, translation = ro.camera_matrix(p_o, R_o, float(model.vis['cam_fovy'][cid]), H, W)
      np.testing.assert_allclose(full[e], image @ focal @ rotation @ translation, atol=1e-9)
  cam = render.Camera(phys, H, W, -1)
  cam.set_pose([0.1, 0.2, 0.3], 2.5, 30.0, -20.0)
  pos, mat = cam.pose()
  p_o, R_o = ro.free_camera_pose([0.1, 0.2, 0.3], 2.5, 30.0, -20.0)
  np.testing.assert_allclose(pos[0].numpy(), p_o, atol=1e-12); np.testing.assert_allclose(mat[0].numpy(), R_o, atol=1e-12)
  image, focal, rotation, translation = ro.camera_matrix(p_o, R_o, 50.0, H, W)
  np.testing.assert_allclose(cam.matrix[0].numpy(), image @ focal @ rotation @ translation, atol=1e-9)
  with pytest.raises(ValueError):
    render.Camera(phys, H, W, 5)
  with pytest.raises(ValueError):
    render.Camera(phys, H, W, 0).set_pose([0, 0, 0], 1, 0, 0)
  # the hand-off to an external renderer: MuJoCo-named host arrays for the chosen environments
  st = render.scene_state(phys, env_ids=[2, 0])
  assert st['geom_xpos'].shape == (2, model.ngeom, 3) and st['geom_xmat'].shape == (2, model.ngeom, 9)
  assert st['cam_xpos'].shape == (2, 5, 3) and st['cam_xmat'].shape == (2, 5, 9)
  p_o, R_o = ro.camera_pose(model.vis, 3, xpos[2], xmat[2], com[2])
  np.testing.assert_allclose(st['cam_xpos'][0, 3], p_o, atol=1e-12); np.testing.assert_allclose(st['cam_xmat'][0, 3].reshape(3, 3), R_o, atol=1e-12)
  np.testing.assert_array_equal(st['geom_type'], model.geom_type); assert st['geom_rgba'].shape == (model.ngeom, 4)


# ---- GPU tier ---------------------------------------------------------------------------------------------------------

EMULATE = os.environ.get('B200MJ_EMULATE_GPU') == '1'
no_renderer_in_emulation = pytest.mark.skipif(EMULATE, reason='the CPU emulation build (tests/emu) holds the physics kernels only: no b200mj_render')


def _gpu_physics(xml, B=2):
  if EMULATE:
    pytest.skip('the CPU emulation build (tests/emu) holds the physics kernels only: no b200mj_render')
  import torch
  from dm_control_b200.physics import BatchedPhysics
  if not torch.cuda.is_available():
    pytest.skip('needs a CUDA device')
  phys = BatchedPhysics(mc.compile_xml(xml), batch=B)
  phys.forward()
  return phys


@pytest.mark.gpu
def test_gpu_depth_render_pin():
  phys = _gpu_physics(PLANE_AND_BOX % DUMMY)
  pixels = phys.render(height=200, width=200, camera_id='top', depth=True)
  assert tuple(pixels.shape) == (2, 200, 200) and str(pixels.dtype) == 'torch.float32'
  for e in range(2):
    np.testing.assert_approx_equal(float(pixels[e].min()), 2.8, 3)
    np.testing.assert_approx_equal(float(pixels[e].max()), 3.0, 3)


@pytest.mark.gpu
def test_gpu_segmentation_render_pin():
  phys = _gpu_physics(BOX_FOUR_CORNERS % DUMMY)
  pixels = phys.render(height=200, width=200, camera_id='top', segmentation=True).cpu().numpy()
  assert pixels.shape == (2, 200, 200, 2)
  for e in range(2):
    _check_segmentation_pin(pixels[e], phys.model)


@pytest.mark.gpu
@pytest.mark.parametrize('camera_id,height,width', [('cam0', 200, 300), (1, 300, 200), (-1, 400, 400)])
def test_gpu_camera_matrix_pin(camera_id, height, width):
  from dm_control_b200 import render
  phys = _gpu_physics(TWO_GEOMS_TWO_CAMERAS % DUMMY)
  camera = render.Camera(phys, width=width, height=height, camera_id=camera_id)
  if camera_id == -1:
    camera.set_pose(*FREE_POSE)
  cam = camera.matrix.cpu().numpy()
  pixels = camera.render(segmentation=True).cpu().numpy()
  gx = phys.data.geom_xpos.reshape(2, -1, 3).cpu().numpy()
  for e in range(2):
    for geom_id in (0, 1):
      xs, ys, s = cam[e].dot(np.array([*gx[e, geom_id], 1.0]))
      row, column = int(round(ys / s)), int(round(xs / s))
      assert tuple(pixels[e, row, column]) == (geom_id, render.OBJ_GEOM)


@pytest.mark.gpu
def test_gpu_render_exceptions():
  from dm_control_b200 import render
  phys = _gpu_physics(TWO_GEOMS_TWO_CAMERAS % DUMMY)
  with pytest.raises(ValueError, match='Only one of depth or segmentation'):      # engine_test.py:327-330
    phys.render(depth=True, segmentation=True)
  with pytest.raises(ValueError):
    render.Camera(phys, camera_id=2)                                              # engine_test.py:304-315
  with pytest.raises(ValueError):
    render.Camera(phys, camera_id=-2)
  with pytest.raises(NotImplementedError):
    phys.render(overlays=('x',))
  img = phys.render(height=48, width=64)                                          # engine_test.py:317-325: the Physics method
  assert tuple(img.shape) == (2, 48, 64, 3) and str(img.dtype) == 'torch.uint8'
  cam = render.Camera(phys, 96, 128)
  cam.set_pose(FREE_POSE[0], 0.3, 90.0, -45.0)                                    # 3 mm pixels: the 5 mm spheres cover some
  before = cam.render().clone()
  assert int(before.max()) > 0
  pose = cam.get_pose()
  cam.set_pose(pose.lookat + np.array([0.01, 0.02, -0.03]), pose.distance * 1.5, pose.azimuth - 15, pose.elevation - 10)
  assert cam.get_pose().distance == pose.distance * 1.5                           # engine_test.py:277-302
  assert not bool((before == cam.render()).all())


def _compare_with_oracle(phys, camera_id, H, W, envs):
  """kernel vs numpy restatement on the engine's own frames: labels equal on all but silhouette pixels, depth to 1e-5."""
  from dm_control_b200 import render
  m, d, B = phys.model, phys.data, phys.batch
  cam = render.Camera(phys, H, W, camera_id)
  seg = cam.render(segmentation=True).cpu().numpy(); dep = cam.render(depth=True).cpu().numpy(); rgb = cam.render().cpu().numpy()
  cp, cm = cam.pose()
  cp, cm = cp.cpu().numpy(), cm.cpu().numpy()
  sizes, stride = cam._sizes()
  sizes = sizes.cpu().numpy()
  gx = d.geom_xpos.reshape(B, -1, 3).cpu().numpy(); gm = d.geom_xmat.reshape(B, -1, 9).cpu().numpy()
  if cam._sites:
    sx = d.site_xpos.reshape(B, -1, 3).cpu().numpy(); sm = d.site_xmat.reshape(B, -1, 9).cpu().numpy()
  else:
    sx, sm = np.zeros((B, 0, 3)), np.zeros((B, 0, 9))
  for e in envs:
    sz = sizes[e] if stride else sizes
    o_rgb, o_dep, o_seg = ro.render(m.vis, np.asarray(m.geom_type), sz[:m.ngeom], gx[e], gm[e], np.asarray(m.site_type), sz[m.ngeom:],
                                    sx[e], sm[e], cp[e], cm[e], cam._fovy, H, W, sites=cam._sites)
    same = (o_seg == seg[e]).all(-1)
    assert same.mean() > 0.998, (e, same.mean())
    assert (o_seg[..., 0] >= 0).mean() > 0.05                                    # the camera does see the model
    np.testing.assert_allclose(dep[e][same], o_dep[same], rtol=2e-6, atol=1e-6)
    assert np.abs(rgb[e][same].astype(int) - o_rgb[same].astype(int)).max() <= 1


@pytest.mark.gpu
@no_renderer_in_emulation
def test_gpu_render_matches_oracle_humanoid():
  import torch
  from dm_control_b200 import suite
  env = suite.load('humanoid', 'run', batch=4, seed=1, outputs='all')
  env.reset()
  a = torch.zeros(4, env.physics.model.nu, dtype=torch.float64, device=env.physics.device)
  for _ in range(3):
    env.step(a)
  phys = env.physics
  ncam = int(phys.model.vis['cam_bodyid'].shape[0])
  assert ncam >= 2                                                               # suite/humanoid.xml: back (trackcom), side (track)
  for cid in range(ncam):
    _compare_with_oracle(phys, cid, 60, 80, range(4))
  _compare_with_oracle(phys, -1, 48, 48, [0])


@pytest.mark.gpu
@no_renderer_in_emulation
def test_gpu_egocentric_camera_in_the_corridor():
  """The CMU walker's head camera (cmu_humanoid.py:448-455) sees that environment's own walls (per-environment geoms)."""
  import torch
  from dm_control_b200 import locomotion
  env = locomotion.load('cmu_humanoid_run_walls', batch=3, seed=2, egocentric_camera=True)
  ts = env.reset()
  img = ts.observation['walker/egocentric_camera']
  assert tuple(img.shape) == (3, 64, 64, 3) and str(img.dtype) == 'torch.uint8'
  assert not bool((img[0] == img[1]).all())                                     # different walls, different images
  phys = env.physics
  _compare_with_oracle(phys, phys.model.name2id('egocentric', 'camera'), 64, 64, range(3))      # (this physics materialises no site frames)
