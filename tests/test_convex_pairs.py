"""The narrow phase for pairs without a closed form (include/b200mj_convex.h: MPR, capsule-box), through the oracle's
test hook. That source is shared with the CUDA engine, so it is held here to answers it did not produce itself: the
closed forms it must reproduce on special cases (an ellipsoid with equal radii is a sphere, ...), and geometry."""
import numpy as np
import pytest

SPHERE, CAPSULE, ELLIPSOID, CYLINDER, BOX = 2, 3, 4, 5, 6
I3 = np.eye(3)


def rot(axis, ang):
  axis = np.asarray(axis, float) / np.linalg.norm(axis)
  K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
  return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


@pytest.fixture(scope='module')
def narrow(oracle_mod):
  return oracle_mod.narrowphase


def test_ball_shaped_ellipsoid_reproduces_sphere_pairs(narrow):
  rs = np.random.RandomState(0)
  for _ in range(40):
    p1, p2 = rs.uniform(-.1, .1, 3), rs.uniform(-.1, .1, 3)
    r1, r2 = rs.uniform(.2, .5, 2)
    R = rot(rs.randn(3), rs.uniform(0, 3))
    ref = narrow(SPHERE, p1, I3, [r1, 0, 0], SPHERE, p2, I3, [r2, 0, 0])
    got = narrow(SPHERE, p1, I3, [r1, 0, 0], ELLIPSOID, p2, R, [r2, r2, r2])
    assert len(ref) == len(got) == 1
    assert abs(ref[0][0] - got[0][0]) < 2e-6
    assert np.abs(ref[0][1] - got[0][1]).max() < 2e-5 and np.abs(ref[0][2] - got[0][2]).max() < 2e-4
    # ellipsoid-ellipsoid as well (both balls)
    got2 = narrow(ELLIPSOID, p1, R.T, [r1, r1, r1], ELLIPSOID, p2, R, [r2, r2, r2])
    assert abs(ref[0][0] - got2[0][0]) < 2e-6 and np.abs(ref[0][2] - got2[0][2]).max() < 2e-4


def test_ball_against_capsule_and_box_matches_closed_forms(narrow):
  rs = np.random.RandomState(1)
  hits = 0
  for _ in range(400):
    R = rot(rs.randn(3), rs.uniform(0, 3))
    pc = rs.uniform(-.2, .2, 3)
    pb = pc + rs.uniform(-.5, .5, 3)
    r = rs.uniform(.15, .3)
    # capsule (type 3) < ellipsoid (type 4): capsule is geom1
    ref = narrow(SPHERE, pb, I3, [r, 0, 0], CAPSULE, pc, R, [.1, .3, 0])
    got = narrow(CAPSULE, pc, R, [.1, .3, 0], ELLIPSOID, pb, I3, [r, r, r])
    assert len(ref) == len(got)
    if ref and ref[0][0] > -0.03:        # shallow contacts only: MPR measures deep overlaps along the centre ray
      hits += 1
      assert -1e-6 <= ref[0][0] - got[0][0] < 0.1 * abs(ref[0][0]) + 2e-4      # never shallower than the true minimum
      assert np.abs(ref[0][2] + got[0][2]).max() < 0.3          # geom order swapped -> normal flips
    # box: sphere-box closed form (centre outside the box) vs ellipsoid-box by MPR
    size = rs.uniform(.1, .3, 3)
    refb = narrow(SPHERE, pb, I3, [r, 0, 0], BOX, pc, R, size)
    gotb = narrow(ELLIPSOID, pb, I3, [r, r, r], BOX, pc, R, size)
    loc = R.T @ (pb - pc)
    if refb and np.any(np.abs(loc) > size) and refb[0][0] > -0.03:
      # MPR measures the overlap close to the centre ray: never shallower than the true minimum, and near it for the
      # shallow contacts that occur in a simulation
      assert len(gotb) == 1 and -1e-6 <= refb[0][0] - gotb[0][0] < 0.1 * abs(refb[0][0]) + 1e-3
      assert np.abs(refb[0][2] - gotb[0][2]).max() < 0.3
    if not refb:
      assert not gotb
  assert hits > 10


def test_separated_shapes_give_no_contact_and_overlap_is_rotation_invariant(narrow):
  rs = np.random.RandomState(2)
  kinds = [(ELLIPSOID, [.3, .2, .1]), (CYLINDER, [.15, .25, 0]), (BOX, [.2, .15, .1]), (CAPSULE, [.1, .2, 0]), (SPHERE, [.2, 0, 0])]
  n_hit = 0
  for _ in range(400):
    (ta, sa), (tb, sb) = kinds[rs.randint(5)], kinds[rs.randint(5)]
    if ta > tb:
      (ta, sa), (tb, sb) = (tb, sb), (ta, sa)
    if ta <= CAPSULE and tb <= CAPSULE:
      continue
    Ra, Rb = rot(rs.randn(3), rs.uniform(0, 3)), rot(rs.randn(3), rs.uniform(0, 3))
    pa, pb = rs.uniform(-.1, .1, 3), rs.uniform(-.5, .5, 3)
    far = pb + 5.0
    assert narrow(ta, pa, Ra, sa, tb, far, Rb, sb) == []
    got = narrow(ta, pa, Ra, sa, tb, pb, Rb, sb)
    if not got:
      continue
    if got[0][0] < -0.04:
      continue          # MPR measures deep overlaps along the centre ray, not the minimum: only shallow contacts are pinned
    n_hit += 1
    Q = rot(rs.randn(3), rs.uniform(0, 3)); shift = rs.uniform(-1, 1, 3)
    got2 = narrow(ta, Q @ pa + shift, Q @ Ra, sa, tb, Q @ pb + shift, Q @ Rb, sb)
    assert len(got2) == len(got)
    for (d, p, n), (d2, p2, n2) in zip(got, got2):
      assert d <= 1e-9 and abs(np.linalg.norm(n) - 1) < 1e-9
      assert abs(d - d2) < 1e-4
      if ta != CAPSULE or tb != BOX:       # MPR: one contact, frame-independent up to its tolerance
        assert np.abs(Q @ n - n2).max() < 0.15 and np.abs(Q @ p + shift - p2).max() < 0.05
  assert n_hit > 20


def test_box_on_box_depth_is_the_face_overlap(narrow):
  got = narrow(BOX, [0, 0, 0], I3, [.5, .5, .1], BOX, [.1, .05, .25], rot([0, 0, 1], .3), [.2, .2, .2])
  assert len(got) == 1
  d, p, n = got[0]
  assert abs(d + 0.05) < 1e-5            # top face z = .1, bottom of the upper box z = .05
  assert np.abs(n - [0, 0, 1]).max() < 1e-3


def test_capsule_on_box_contact_pattern(narrow):
  box_p, box_s = [0, 0, 0], [.5, .4, .1]
  # lying flat on the top face: two contacts at the ends, equal depth
  flat = narrow(CAPSULE, [0, 0, .1 + .05 - .01], rot([0, 1, 0], np.pi / 2), [.05, .2, 0], BOX, box_p, I3, box_s)
  assert len(flat) == 2
  for d, p, n in flat:
    assert abs(d + 0.01) < 1e-9 and np.abs(n - [0, 0, -1]).max() < 1e-9
  assert abs(abs(flat[0][1][0] - flat[1][1][0]) - 0.4) < 1e-9
  # inclined: only the lower end touches
  tilt = narrow(CAPSULE, [0, 0, .1 + .05 + .1 * np.sin(.5) - .005], rot([0, 1, 0], np.pi / 2 - .5), [.05, .1, 0], BOX, box_p, I3, box_s)
  assert len(tilt) == 1 and abs(tilt[0][0] + .005) < 1e-6
  # across the edge x = .5 of the top face, tilted down outside: the contact sits at the edge, inside the segment
  cross = narrow(CAPSULE, [.5, 0, .1 + .03], rot([0, 1, 0], np.pi / 2 + .4), [.05, .3, 0], BOX, box_p, I3, box_s)
  assert len(cross) >= 1
  assert any(abs(p[0] - .5) < .06 for d, p, n in cross)
  # clear of the box
  assert narrow(CAPSULE, [0, 0, .5], I3, [.05, .2, 0], BOX, box_p, I3, box_s) == []


# ---- the engine's warp-cooperative form of cvx_pair (b200mj.cu: cvx_pair_warp) ------------------------------------------
_WARP_CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests/emu')
import b200mj_emu as emu
from dm_control_b200 import mjcf_compile as mc
from oracle import oracle as om
om.build()
TYPES = [('ellipsoid', '.3 .2 .15'), ('cylinder', '.2 .25'), ('box', '.25 .2 .15'), ('capsule', '.15 .25'), ('sphere', '.2')]
rs = np.random.RandomState(%(seed)d)
worst, ncontact, nmis = 0.0, 0, 0
for ta, sa in TYPES[:3]:
  for tb, sb in TYPES:
    xml = f"""<mujoco><option gravity="0 0 0"/><worldbody>
      <body name="a" pos="0 0 0"><freejoint/><geom name="ga" type="{ta}" size="{sa}"/></body>
      <body name="b" pos="0.3 0 0"><freejoint/><geom name="gb" type="{tb}" size="{sb}"/></body>
    </worldbody></mujoco>"""
    m = mc.compile_xml(xml)
    B = 8
    p = emu.EmuPhysics(m, B)
    q = np.zeros((B, 14))
    for e in range(B):
      qa = rs.randn(4); qa /= np.linalg.norm(qa); qb = rs.randn(4); qb /= np.linalg.norm(qb)
      off = rs.randn(3); off *= rs.uniform(0.1, 0.45) / np.linalg.norm(off)
      q[e] = np.concatenate([[0, 0, 0], qa, off, qb])
    p.data.qpos[:] = q; p.forward()
    for e in range(B):
      o = om.OraclePhysics(m); o.qpos[:] = q[e]; o.forward()
      if int(p.data.ncon[e]) != o.ncon:
        nmis += 1; continue
      for k in range(o.ncon):
        cd = o.contact[k]
        worst = max(worst, abs(p.data.contact_dist[e, k] - cd.dist), float(np.abs(p.data.contact_pos[e, k] - np.asarray(cd.pos)).max()),
                    float(np.abs(p.data.contact_frame[e, k][:3] - np.asarray(cd.frame)[:3]).max()))
        ncontact += 1
print(json.dumps(dict(worst=worst, contacts=ncontact, mismatches=nmis)))
'''


@pytest.mark.timeout(900)
@pytest.mark.parametrize('warp', ['1', '0'])
def test_engine_convex_contacts_equal_the_oracles_bitwise(warp):
  """15 type pairs x 8 random poses through the kernel source on the CPU emulation: the warp-cooperative cvx_pair_warp
  (B200MJ_CVX_WARP=1, default) and the one-pair-per-lane cvx_pair give the oracle's contact EXACTLY (same arithmetic,
  same selection order) — distance, position and normal to the last bit."""
  import json, os, subprocess, sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, '-c', _WARP_CHILD % dict(root=root, seed=5)], env=dict(os.environ, B200MJ_CVX_WARP=warp),
                     capture_output=True, text=True, timeout=800)
  assert r.returncode == 0, r.stderr[-2000:]
  out = json.loads(r.stdout.strip().splitlines()[-1])
  assert out['contacts'] > 80 and out['mismatches'] == 0 and out['worst'] == 0.0, out
