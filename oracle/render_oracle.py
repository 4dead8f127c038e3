"""CPU oracle of the rendering hand-off (TEST INFRASTRUCTURE ONLY — never imported by dm_control_b200/).

Restates, in numpy and one environment at a time, what `dm_control_b200/csrc/b200mj_render.cu` computes on the device:

  * the camera model of the reference's `engine.Camera.matrices` (dm_control/mujoco/engine.py:759-810): the camera looks
    along -z of its frame, +x is right and +y is up; pixel (row v, column u) sees the ray
        d_cam = ((u - cx) / f, -(v - cy) / f, -1),   f = (height / 2) / tan(fovy / 2),  cx = (W-1)/2, cy = (H-1)/2,
    which is exactly the inverse of `image @ focal @ rotation @ translation`;
  * MuJoCo's camera modes (mjtCamLight: fixed, track, trackcom, targetbody[com]) for `cam_xpos / cam_xmat`;
  * ray / primitive intersections with the semantics of MuJoCo's mj_ray family (plane: front face only, finite when
    size > 0; sphere; capsule; ellipsoid; cylinder; box), nearest hit beyond the near plane;
  * the three images of `Physics.render` (engine.py:178-233, 840-946): flat-shaded RGB (a headlight; NOT pixel-identical
    to MuJoCo's OpenGL renderer: no textures, shadows, reflections or skybox), `depth` = distance along the optical axis
    (what the reference derives from the z-buffer, engine.py:917-924: orthographic depth, background at the far plane), and
    `segmentation` = (object id, object type) per pixel with (-1, -1) for the background (engine.py:926-944).

Pinned by the reference's own known-answers in tests/test_render.py (engine_test.py:64-131 depth / segmentation,
:228-275 camera matrix).
"""
from __future__ import annotations

import numpy as np

OBJ_GEOM, OBJ_SITE = 5, 6           # mjtObj
ZNEAR, ZFAR = 0.01, 50.0            # mjModel.vis.map defaults, in units of stat.extent
AMBIENT, DIFFUSE = 0.4, 0.6         # headlight shading of the RGB image
TMIN_EPS = 1e-9


def quat_to_mat(q):
  w, x, y, z = q
  return np.array([[w*w + x*x - y*y - z*z, 2*(x*y - w*z), 2*(x*z + w*y)],
                   [2*(x*y + w*z), w*w - x*x + y*y - z*z, 2*(y*z - w*x)],
                   [2*(x*z - w*y), 2*(y*z + w*x), w*w - x*x - y*y + z*z]])


def look_at(cam_pos, target):
  """Frame whose -z axis points at `target`, x horizontal (MuJoCo's targetbody cameras)."""
  z = cam_pos - target
  z = z / max(np.linalg.norm(z), 1e-12)
  x = np.cross(np.array([0.0, 0.0, 1.0]), z)
  if np.linalg.norm(x) < 1e-9:
    x = np.array([1.0, 0.0, 0.0])
  x = x / np.linalg.norm(x)
  y = np.cross(z, x)
  return np.stack([x, y, z], axis=1)


def camera_pose(vis, cam_id, xpos, xmat, subtree_com):
  """cam_xpos [3], cam_xmat [3,3] of a model camera from body frames (one environment)."""
  b, mode = int(vis['cam_bodyid'][cam_id]), int(vis['cam_mode'][cam_id])
  R = np.asarray(xmat[b]).reshape(3, 3)
  if mode == 0 or mode >= 3:
    pos = xpos[b] + R @ vis['cam_pos'][cam_id]
    mat = R @ quat_to_mat(vis['cam_quat'][cam_id])
    if mode >= 3:
      t = int(vis['cam_targetbodyid'][cam_id])
      mat = look_at(pos, subtree_com[t] if mode == 4 else xpos[t])
  elif mode == 1:
    pos, mat = xpos[b] + vis['cam_pos0'][cam_id], vis['cam_mat0'][cam_id].reshape(3, 3)
  else:
    pos, mat = subtree_com[b] + vis['cam_poscom0'][cam_id], vis['cam_mat0'][cam_id].reshape(3, 3)
  return pos, mat


def free_camera_pose(lookat, distance, azimuth, elevation):
  """mjvCamera (free): the camera sits `distance` behind `lookat` along the azimuth / elevation direction."""
  az, el = np.deg2rad(azimuth), np.deg2rad(elevation)
  forward = np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
  pos = np.asarray(lookat) - distance * forward
  return pos, look_at(pos, np.asarray(lookat))


def camera_matrix(cam_xpos, cam_xmat, fovy, height, width):
  """image @ focal @ rotation @ translation, engine.py:784-810."""
  translation = np.eye(4); translation[:3, 3] = -cam_xpos
  rotation = np.eye(4); rotation[:3, :3] = np.asarray(cam_xmat).reshape(3, 3).T
  f = (1.0 / np.tan(np.deg2rad(fovy) / 2)) * height / 2.0
  focal = np.diag([-f, f, 1.0, 0])[0:3, :]
  image = np.eye(3); image[0, 2] = (width - 1) / 2.0; image[1, 2] = (height - 1) / 2.0
  return image, focal, rotation, translation


def _pick(t1, t2, tmin):
  """nearest of two roots that lies beyond tmin (inf if none)."""
  a = np.where(t1 > tmin, t1, np.inf)
  b = np.where(t2 > tmin, t2, np.inf)
  return np.minimum(a, b)


def _ray_sphere(o, d, r, tmin, center=None):
  oc = o if center is None else o - center
  a = np.sum(d * d, -1); b = np.sum(oc * d, -1); c = np.sum(oc * oc, -1) - r * r
  disc = b * b - a * c
  ok = disc >= 0
  sq = np.sqrt(np.where(ok, disc, 0.0))
  t = _pick((-b - sq) / a, (-b + sq) / a, tmin)
  return np.where(ok, t, np.inf)


def ray_primitive(gtype, size, o, d, tmin):
  """Rays (o, d) [N,3] in the primitive's frame -> (t [N] (inf = miss), outward normal [N,3] in that frame)."""
  N = o.shape[0]
  t = np.full(N, np.inf); nrm = np.zeros((N, 3))
  if gtype == 0:      # plane z = 0, front face only, finite where size > 0
    with np.errstate(divide='ignore', invalid='ignore'):
      tt = -o[:, 2] / d[:, 2]
    p = o + tt[:, None] * d
    ok = (d[:, 2] < -1e-15) & (tt > tmin)
    if size[0] > 0: ok &= np.abs(p[:, 0]) <= size[0]
    if size[1] > 0: ok &= np.abs(p[:, 1]) <= size[1]
    t = np.where(ok, tt, np.inf); nrm[:, 2] = 1.0
  elif gtype == 2:
    t = _ray_sphere(o, d, size[0], tmin)
    nrm = o + np.where(np.isfinite(t), t, 0.0)[:, None] * d
  elif gtype == 4:    # ellipsoid: unit sphere in scaled coordinates
    s = np.asarray(size, dtype=np.float64)
    t = _ray_sphere(o / s, d / s, 1.0, tmin)
    nrm = (o + np.where(np.isfinite(t), t, 0.0)[:, None] * d) / (s * s)
  elif gtype in (3, 5):   # capsule / cylinder: radius size[0], half-length size[1] along z
    r, h = size[0], size[1]
    a = d[:, 0]**2 + d[:, 1]**2; b = o[:, 0] * d[:, 0] + o[:, 1] * d[:, 1]; c = o[:, 0]**2 + o[:, 1]**2 - r * r
    disc = b * b - a * c
    ok = (disc >= 0) & (a > 1e-30)
    sq = np.sqrt(np.where(ok, disc, 0.0))
    with np.errstate(divide='ignore', invalid='ignore'):
      t1, t2 = (-b - sq) / a, (-b + sq) / a
    z1, z2 = o[:, 2] + t1 * d[:, 2], o[:, 2] + t2 * d[:, 2]
    t1 = np.where(ok & (t1 > tmin) & (np.abs(z1) <= h), t1, np.inf)
    t2 = np.where(ok & (t2 > tmin) & (np.abs(z2) <= h), t2, np.inf)
    ts = np.minimum(t1, t2)
    ps = o + np.where(np.isfinite(ts), ts, 0.0)[:, None] * d
    ns = np.stack([ps[:, 0], ps[:, 1], np.zeros(N)], -1)
    t, nrm = ts, ns
    for sgn in (1.0, -1.0):
      if gtype == 3:    # hemispherical cap: the part of the sphere beyond the cylinder's end
        cen = np.array([0.0, 0.0, sgn * h])
        oc = o - cen
        aa = np.sum(d * d, -1); bb = np.sum(oc * d, -1); cc = np.sum(oc * oc, -1) - r * r
        dd = bb * bb - aa * cc
        okc = dd >= 0
        sq = np.sqrt(np.where(okc, dd, 0.0))
        for tc in ((-bb - sq) / aa, (-bb + sq) / aa):
          zc = o[:, 2] + tc * d[:, 2]
          good = okc & (tc > tmin) & (sgn * zc >= h) & (tc < t)
          pc = o + np.where(good, tc, 0.0)[:, None] * d
          t = np.where(good, tc, t); nrm = np.where(good[:, None], pc - cen, nrm)
      else:             # flat cap
        with np.errstate(divide='ignore', invalid='ignore'):
          tc = (sgn * h - o[:, 2]) / d[:, 2]
        pc = o + np.where(np.isfinite(tc), tc, 0.0)[:, None] * d
        good = np.isfinite(tc) & (tc > tmin) & (pc[:, 0]**2 + pc[:, 1]**2 <= r * r) & (tc < t)
        t = np.where(good, tc, t); nrm = np.where(good[:, None], np.array([0.0, 0.0, sgn]), nrm)
  elif gtype == 6:    # box: slabs
    s = np.asarray(size, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
      inv = 1.0 / d
      ta, tb = (-s - o) * inv, (s - o) * inv
    par = np.abs(d) < 1e-300
    inside = np.abs(o) <= s
    lo = np.where(par, np.where(inside, -np.inf, np.inf), np.minimum(ta, tb))
    hi = np.where(par, np.where(inside, np.inf, -np.inf), np.maximum(ta, tb))
    tn, tf = lo.max(-1), hi.min(-1)
    hit = tn <= tf
    tt = np.where(tn > tmin, tn, tf)
    ok = hit & (tt > tmin)
    t = np.where(ok, tt, np.inf)
    p = o + np.where(ok, tt, 0.0)[:, None] * d
    ax = np.argmax(np.abs(p) / s, -1)
    nrm = np.zeros((N, 3)); nrm[np.arange(N), ax] = np.sign(p[np.arange(N), ax])
  else:
    raise NotImplementedError(gtype)
  return t, nrm


def pixel_rays(cam_xmat, fovy, height, width):
  f = (height / 2.0) / np.tan(np.deg2rad(fovy) / 2)
  v, u = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing='ij')
  dc = np.stack([(u - (width - 1) / 2.0) / f, -(v - (height - 1) / 2.0) / f, -np.ones_like(u)], -1).reshape(-1, 3)
  return dc @ np.asarray(cam_xmat).reshape(3, 3).T      # world directions; depth = ray parameter


def render(vis, geom_type, geom_size, geom_xpos, geom_xmat, site_type, site_size, site_xpos, site_xmat, cam_xpos, cam_xmat,
           fovy, height, width, sites=True, groups=(0, 1, 2)):
  """-> rgb uint8 [H,W,3], depth float64 [H,W], seg int32 [H,W,2] for one environment."""
  extent = float(vis['stat_extent'][0])
  near, far = ZNEAR * extent, ZFAR * extent
  d = pixel_rays(cam_xmat, fovy, height, width)
  o = np.broadcast_to(np.asarray(cam_xpos, dtype=np.float64), d.shape)
  N = d.shape[0]
  best = np.full(N, far); seg = np.full((N, 2), -1, np.int32); col = np.zeros((N, 3))
  objs = [(OBJ_GEOM, i, int(geom_type[i]), geom_size[i], geom_xpos[i], geom_xmat[i], vis['geom_rgba'][i], int(vis['geom_group'][i]))
          for i in range(len(geom_type))]
  if sites:
    objs += [(OBJ_SITE, i, int(site_type[i]), site_size[i], site_xpos[i], site_xmat[i], vis['site_rgba'][i], int(vis['site_group'][i]))
             for i in range(len(site_type))]
  for kind, i, typ, size, pos, mat, rgba, group in objs:
    if group not in groups or rgba[3] == 0:
      continue
    R = np.asarray(mat).reshape(3, 3)
    ol, dl = (o - pos) @ R, d @ R
    t, nl = ray_primitive(typ, np.asarray(size, dtype=np.float64), ol, dl, near)
    hit = t < best
    if not hit.any():
      continue
    n = nl @ R.T
    n = n / np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-300)
    dn = d / np.linalg.norm(d, axis=-1, keepdims=True)
    shade = AMBIENT + DIFFUSE * np.maximum(0.0, -np.sum(n * dn, -1))
    c = np.asarray(rgba[:3], dtype=np.float64)[None, :] * shade[:, None]
    best = np.where(hit, t, best)
    seg[hit] = (i, kind)
    col = np.where(hit[:, None], c, col)
  rgb = np.clip(np.floor(col * 255.0 + 0.5), 0, 255).astype(np.uint8)
  return rgb.reshape(height, width, 3), best.reshape(height, width), seg.reshape(height, width, 2)
