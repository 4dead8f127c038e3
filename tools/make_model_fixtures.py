"""Compile the reference's MJCF models (read from /root/reference, THIS container only) into model fixtures.

/root/reference does not exist on the GPU box, and reference sources must not be copied into this repo, so the
hot-path configs travel as *compiled* tables (dm_control_b200/assets/<name>.npz), produced by this repo's own
MJCF compiler from the reference XML:

  cartpole   dm_control/suite/cartpole.xml                         (suite.cartpole:swingup)
  cheetah    dm_control/suite/cheetah.xml                          (suite.cheetah:run)
  humanoid   dm_control/suite/humanoid.xml                         (suite.humanoid:run)
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


if __name__ == '__main__':
  main()
