"""Batched twins of the hot-path suite domains (reference: dm_control/suite/{cartpole,cheetah,humanoid,lqr,quadruped}.py).

`load(domain, task, batch=...)` mirrors `suite.load` (suite/__init__.py:93-114) and returns a
`control.BatchedEnvironment`.
"""
from __future__ import annotations

from . import cartpole, cheetah, humanoid, lqr, quadruped

_DOMAINS = dict(cartpole=cartpole, cheetah=cheetah, humanoid=humanoid, lqr=lqr, quadruped=quadruped)


def load(domain_name, task_name, batch=1, seed=0, **kw):
  if domain_name not in _DOMAINS:
    raise ValueError(f'Domain {domain_name!r} does not exist in the batched suite (have {sorted(_DOMAINS)}).')
  mod = _DOMAINS[domain_name]
  if task_name not in mod.TASKS:
    raise ValueError(f'Level {task_name!r} does not exist in domain {domain_name!r}.')
  return mod.TASKS[task_name](batch=batch, seed=seed, **kw)
