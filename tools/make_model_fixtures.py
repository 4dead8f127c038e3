"""Compile the reference's MJCF models (read from /root/reference, THIS container only) into model fixtures.

/root/reference does not exist on the GPU box, and reference sources must not be copied into this repo, so the
hot-path configs travel as *compiled* tables (dm_control_b200/assets/<name>.npz), produced by this repo's own
MJCF compiler from the reference XML:

  cartpole   dm_control/suite/cartpole.xml                         (suite.cartpole:swingup)
  cheetah    dm_control/suite/cheetah.xml                          (suite.cheetah:run)
  humanoid   dm_control/suite/humanoid.xml                         (suite.humanoid:run)
  cmu_humanoid  locomotion/walkers/assets/humanoid_CMU_V2019.xml + the position actuators cmu_humanoid.py builds, on a plane
  quadruped  dm_control/suite/quadruped.xml, stripped exactly as suite/quadruped.py:55-93 `make_model` does for
             `walk` (walls, ball, target site, terrain hfield and rangefinder sensors removed; floor resized)

Run:  python tools/make_model_fixtures.py
"""
import os, sys
import xml.etree.ElementTree as ET
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dm_control_b200 import mjcf_compile as mc

REF = '/root/reference/dm_control/suite'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dm_control_b200', 'assets')


def quadruped_walk_xml():
  root = ET.parse(os.path.join(REF, 'quadruped.xml')).getroot()
  floor_size = 20 * 0.5   # _DEFAULT_TIME_LIMIT * _WALK_SPEED  (suite/quadruped.py:29,41,98)
  parent = {c: p for p in root.iter() for c in p}
  def drop(pred):
    for el in list(root.iter()):
      if pred(el):
        parent[el].remove(el)
  for g in root.iter('geom'):
    if g.get('name') == 'floor':
      g.set('size', f'{floor_size} {floor_size} .5')
  drop(lambda e: e.tag == 'geom' and e.get('name') in ('wall_px', 'wall_py', 'wall_nx', 'wall_ny', 'terrain'))
  drop(lambda e: e.tag == 'body' and e.get('name') == 'ball')
  drop(lambda e: e.tag == 'site' and e.get('name') == 'target')
  drop(lambda e: e.tag == 'rangefinder')
  drop(lambda e: e.tag == 'hfield')
  return ET.tostring(root)


def cmu_humanoid_flat_xml():
  """CMU humanoid V2019 as `CMUHumanoidPositionControlled` builds it (locomotion/walkers/cmu_humanoid.py:360-394,
  scaled_actuators.py:37-82), standing on a plane, under the composer root compiler settings
  (composer/arena.xml:2: angle=radian, boundmass=1e-5, boundinertia=1e-11) and the example's 5 ms physics step
  (locomotion/examples/basic_cmu_2019.py:57-58). Corridor walls and the egocentric camera are NOT part of this fixture."""
  import re
  walkers = '/root/reference/dm_control/locomotion/walkers'
  root = ET.parse(os.path.join(walkers, 'assets', 'humanoid_CMU_V2019.xml')).getroot()
  src = open(os.path.join(walkers, 'cmu_humanoid.py')).read()
  table = src.split('_POSITION_ACTUATORS = [')[1].split(']\n')[0]
  params = re.findall(r"PositionActuatorParams\('(\w+)',\s*\[\s*(-?[\d.]+),\s*(-?[\d.]+)\s*\],\s*([\d.]+)\s*\)", table)
  assert len(params) == 56, len(params)
  root.insert(0, ET.Element('compiler', angle='radian', boundmass='1e-5', boundinertia='1e-11'))
  root.insert(1, ET.Element('option', timestep='0.005'))
  world = root.find('worldbody')
  world.insert(0, ET.Element('geom', name='groundplane', type='plane', size='1 1 1', condim='3', friction='1 0.005 0.0001',
                             solref='0.02 1', solimp='0.9 0.95 0.001 0.5 2', contype='1', conaffinity='1'))
  body = [b for b in world.findall('body') if b.get('name') == 'root'][0]
  body.set('pos', '0 0 0.94'); body.set('quat', '0.859 1.0 1.0 0.859')          # cmu_humanoid.py:174-176 upright pose
  body.insert(0, ET.Element('freejoint', name='root'))
  joints = {j.get('name'): j for j in root.iter('joint')}
  # joint ranges: explicit on the element, else inherited from defaults (all CMU joints carry their own range)
  act = root.find('actuator')
  for m in list(act):
    act.remove(m)
  for name, f_lo, f_hi, kp in params:
    lo, hi = [float(x) for x in joints[name].get('range').split()]
    kp = float(kp); slope = (hi - lo) / 2.0
    ET.SubElement(act, 'general', name=name, joint=name, biastype='affine', gainprm=repr(kp * slope),
                  biasprm=f'{kp * (lo - slope * -1.0)!r} {-kp!r} 0', ctrllimited='true', ctrlrange='-1 1',
                  forcelimited='true', forcerange=f'{f_lo} {f_hi}')
  return ET.tostring(root)


N_WALLS = 25      # basic_cmu_2019.py:42-48: wall_gap 4, corridor_length 100, no initial padding -> walls at x = 2, 6, ..., 98


def cmu_corridor_walls_xml():
  """`cmu_humanoid_run_walls` (locomotion/examples/basic_cmu_2019.py:34-63) as ONE model: the WallsCorridor arena
  (arenas/corridors.py:94-120: ground plane + four side planes; :394-440: wall boxes) with the position-controlled CMU
  humanoid attached by a free joint, and the four `framepos` end-effector sensors the walker adds when it is built
  (walkers/legacy_base.py:222-233: objtype xbody, reftype xbody = root). The 25 wall boxes get PLACEHOLDER pos / size:
  every environment overrides them per episode (BatchedPhysics.set_variable_geoms), which stands where the
  reference's per-episode recompile stands (composer/environment.py:378-383). The egocentric camera observable is not
  part of the batched task (rendering is out of scope)."""
  root = ET.fromstring(cmu_humanoid_flat_xml())
  world = root.find('worldbody')
  world.remove([g for g in world.findall('geom') if g.get('name') == 'groundplane'][0])
  L, W, PAD, SH = 100.0, 10.0, 2.0, 4.0      # corridor_length, corridor_width, _CORRIDOR_X_PADDING, _SIDE_WALL_HEIGHT
  # the arena's geoms live in the arena's own default scope under composer (MuJoCo built-in defaults), not in the walker's
  # (`<default><geom condim="1" friction=".7" solref=".015 1" .../>`, humanoid_CMU_V2019.xml:8-9): spell them out
  builtin = dict(condim='3', friction='1 0.005 0.0001', solref='0.02 1', solimp='0.9 0.95 0.001 0.5 2', contype='1', conaffinity='1')
  def plane(name, pos, size, xyaxes=None):
    a = dict(name=name, type='plane', pos=' '.join(map(str, pos)), size=' '.join(map(str, size)), **builtin)
    if xyaxes:
      a['xyaxes'] = xyaxes
    return ET.Element('geom', **a)
  arena = [plane('ground_plane', (L / 2, 0, 0), (L / 2 + PAD, W / 2, 1)),
           plane('left_plane', (L / 2, W / 2, SH / 2), (L / 2 + PAD, SH / 2, 1), '1 0 0 0 0 1'),
           plane('right_plane', (L / 2, -W / 2, SH / 2), (L / 2 + PAD, SH / 2, 1), '-1 0 0 0 0 1'),
           plane('near_plane', (-PAD, 0, SH / 2), (W / 2, SH / 2, 1), '0 1 0 0 0 1'),
           plane('far_plane', (L + PAD, 0, SH / 2), (W / 2, SH / 2, 1), '0 -1 0 0 0 1')]
  for k, g in enumerate(arena):
    world.insert(k, g)
  walls = ET.Element('body', name='walls')
  for k in range(N_WALLS):
    side = 1 if k % 2 == 0 else -1
    ET.SubElement(walls, 'geom', name=f'wall_{k}', type='box', pos=f'{2.0 + 4.0 * k} {side * 3.0} 1.5', size='0.08 2.0 1.5', **builtin)
  world.insert(len(arena), walls)
  body = [b for b in world.findall('body') if b.get('name') == 'root'][0]
  body.set('pos', '0.5 0 0.94')          # walker_spawn_position (0.5, 0, 0) on top of the upright pose
  sensor = root.find('sensor')
  for eff in ('rradius', 'lradius', 'rfoot', 'lfoot'):     # cmu_humanoid.py:330-335 end_effectors
    ET.SubElement(sensor, 'framepos', name=f'{eff}_end_effector', objtype='xbody', objname=eff, reftype='xbody', refname='root')
  return ET.tostring(root)


def main():
  os.makedirs(OUT, exist_ok=True)
  caps = dict(cartpole=dict(nconmax=0, njmax=4), cheetah=dict(nconmax=16, njmax=80),
              humanoid=dict(nconmax=32, njmax=96), quadruped=dict(nconmax=24, njmax=96))
  for name in ('cartpole', 'cheetah', 'humanoid'):
    m = mc.compile_file(os.path.join(REF, name + '.xml'), **caps[name])
    m.save(os.path.join(OUT, name + '.npz'))
    print(name, 'nq', m.nq, 'nv', m.nv, 'nu', m.nu, 'nbody', m.nbody, 'ngeom', m.ngeom, 'npair', m.npair, 'nconmax', m.nconmax, 'njmax', m.njmax)
  try:
    m = mc.compile_xml(quadruped_walk_xml(), base_dir=REF, **caps['quadruped'])
    m.save(os.path.join(OUT, 'quadruped.npz'))
    print('quadruped', 'nq', m.nq, 'nv', m.nv, 'nu', m.nu, 'na', m.na, 'nbody', m.nbody, 'ngeom', m.ngeom, 'npair', m.npair)
  except Exception as ex:
    print('quadruped: not compiled yet:', repr(ex))
  m = mc.compile_xml(cmu_corridor_walls_xml(), nconmax=48, njmax=200)
  m.save(os.path.join(OUT, 'cmu_corridor_walls.npz'))
  print('cmu_corridor_walls', 'nq', m.nq, 'nv', m.nv, 'nu', m.nu, 'nbody', m.nbody, 'ngeom', m.ngeom, 'npair', m.npair, 'nsensordata', m.nsensordata)
  m = mc.compile_xml(cmu_humanoid_flat_xml(), nconmax=40, njmax=112)
  m.save(os.path.join(OUT, 'cmu_humanoid.npz'))
  print('cmu_humanoid', 'nq', m.nq, 'nv', m.nv, 'nu', m.nu, 'nbody', m.nbody, 'ngeom', m.ngeom, 'npair', m.npair, 'nsensordata', m.nsensordata,
        'mass %.2f' % m.body_mass.sum())


if __name__ == '__main__':
  main()
