"""N>1 host logic on CPU: world_size-2 gloo processes exercise the shard map and the observation gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dm_control_b200 import sharding


def test_shard_range_partitions_exactly():
  for B in (1, 7, 8, 8192, 8191):
    for W in (1, 2, 3, 4, 8):
      spans = [sharding.shard_range(B, W, r) for r in range(W)]
      assert spans[0][0] == 0 and spans[-1][1] == B
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      sizes = [b - a for a, b in spans]
      assert max(sizes) - min(sizes) <= 1
      assert sizes == sharding.shard_sizes(B, W)
  with pytest.raises(ValueError):
    sharding.shard_range(8, 2, 2)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, global_batch, ret):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  r, w, _ = sharding.init_from_env(backend='gloo')
  lo, hi = sharding.shard_range(global_batch, w, r)
  # each environment's packed [obs(67), reward, discount] row is a function of its GLOBAL index only
  idx = torch.arange(lo, hi, dtype=torch.float64)
  block = idx[:, None] * 1000 + torch.arange(69, dtype=torch.float64)[None, :]
  for step in range(3):
    got = sharding.gather_to_rank0(block + step, global_batch)
    if r == 0:
      want = torch.arange(global_batch, dtype=torch.float64)[:, None] * 1000 + torch.arange(69, dtype=torch.float64)[None, :] + step
      ret[step] = bool(torch.equal(got, want))
    else:
      assert got is None
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.parametrize('global_batch', [64, 65])
def test_gather_to_rank0_world2_gloo(global_batch):
  ctx = mp.get_context('spawn')
  ret = ctx.Manager().dict()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, ret)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(120)
    assert p.exitcode == 0
  assert dict(ret) == {0: True, 1: True, 2: True}


def _worker_allgather(rank, world, port, n_local, ret):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
  r, w, _ = sharding.init_from_env(backend='gloo')
  idx = torch.arange(r * n_local, (r + 1) * n_local, dtype=torch.float64)
  block = idx[:, None] * 1000 + torch.arange(69, dtype=torch.float64)[None, :]
  out = sharding.alloc_gather(block, w)
  ok = True
  for step in range(3):
    sharding.gather_packed(block + step, out)       # same preallocated destination every step
    want = torch.arange(w * n_local, dtype=torch.float64)[:, None] * 1000 + torch.arange(69, dtype=torch.float64)[None, :] + step
    ok = ok and bool(torch.equal(out, want))
  ret[r] = ok
  dist.barrier()
  dist.destroy_process_group()


def test_gather_packed_world2_gloo():
  """bench.py's per-step exchange: one all_gather_into_tensor into a preallocated block, environment order on every rank."""
  ctx = mp.get_context('spawn')
  ret = ctx.Manager().dict()
  port = _free_port()
  procs = [ctx.Process(target=_worker_allgather, args=(r, 2, port, 32, ret)) for r in range(2)]
  for p in procs:
    p.start()
  for p in procs:
    p.join(120)
    assert p.exitcode == 0
  assert dict(ret) == {0: True, 1: True}


def test_gather_packed_single_process_is_a_copy():
  block = torch.arange(12, dtype=torch.float64).reshape(4, 3)
  out = sharding.alloc_gather(block, 1)
  assert torch.equal(sharding.gather_packed(block, out), block)
