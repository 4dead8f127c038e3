"""Batched twin of BASELINE.json config 5: the composer CMU-humanoid run-through-corridor task.

Reference pieces (all under dm_control/): `locomotion/examples/basic_cmu_2019.py:34-63` (the environment),
`composer/environment.py:412-465` (step / substep / hook order), `locomotion/tasks/corridors.py:33-158` (task),
`locomotion/arenas/corridors.py:94-175,330-440` (arena + per-episode walls), `locomotion/walkers/cmu_humanoid.py`,
`legacy_base.py`, `base.py` (walker, observables). `egocentric_camera=True` adds the walker's 64 x 64 head-camera
observable, ray-cast on the device by the rendering hand-off (dm_control_b200/render.py; not MuJoCo's OpenGL pixels).
"""
from __future__ import annotations

from . import composer, corridors


def load(name, batch=1, seed=0, **kw):
  """`basic_cmu_2019.cmu_humanoid_run_walls()` for `batch` environments."""
  if name != 'cmu_humanoid_run_walls':
    raise ValueError(f'{name!r}: the batched locomotion suite has cmu_humanoid_run_walls only')
  return corridors.cmu_humanoid_run_walls(batch=batch, seed=seed, **kw)
