"""Environment sharding across the GPUs of one box: one process per GPU, no data-path collective.

Environments are independent, so rank r owns the contiguous range `shard_range(global_batch, world, r)` for the whole
rollout (state never leaves its GPU). The only exchange is the one BASELINE.json's north_star names: per control step
the packed `[n_local, obs_dim + 2]` (observation, reward, discount) block is gathered to rank 0 (`gather_to_rank0`,
NCCL over NVLink on GPUs; the same code runs on gloo/CPU in tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(global_batch, world_size, rank):
  """Contiguous, balanced: the first `global_batch % world_size` ranks get one extra environment."""
  if not 0 <= rank < world_size:
    raise ValueError(f'rank {rank} outside world of {world_size}')
  base, extra = divmod(global_batch, world_size)
  start = rank * base + min(rank, extra)
  return start, start + base + (1 if rank < extra else 0)


def shard_sizes(global_batch, world_size):
  return [shard_range(global_batch, world_size, r)[1] - shard_range(global_batch, world_size, r)[0] for r in range(world_size)]


def init_from_env(backend=None):
  """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun) and initialises the process group if needed."""
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if world > 1 and not dist.is_initialized():
    backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
    kw = {}
    if backend == 'nccl':
      torch.cuda.set_device(local)
      kw['device_id'] = torch.device('cuda', local)
    dist.init_process_group(backend, **kw)
  return rank, world, local


def gather_to_rank0(block, global_batch, out=None):
  """Gather every rank's `[n_local, k]` block to rank 0 in environment order. Returns `[global_batch, k]` on rank 0
  (written into `out` when given), None elsewhere. Shards may differ in size by one row."""
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return block
  world, rank = dist.get_world_size(), dist.get_rank()
  sizes = shard_sizes(global_batch, world)
  if block.shape[0] != sizes[rank]:
    raise ValueError(f'rank {rank}: block has {block.shape[0]} rows, shard has {sizes[rank]}')
  pad = max(sizes)
  send = block if block.shape[0] == pad else torch.cat([block, block.new_zeros(pad - block.shape[0], block.shape[1])])
  if rank == 0:
    bufs = [torch.empty_like(send) for _ in range(world)]
    dist.gather(send, bufs, dst=0)
    res = out if out is not None else block.new_empty(global_batch, block.shape[1])
    at = 0
    for r, n in enumerate(sizes):
      res[at:at + n] = bufs[r][:n]
      at += n
    return res
  dist.gather(send, None, dst=0)
  return None


def alloc_gather(block, world_size=None):
  """Preallocated `[world * n_local, k]` destination of `gather_packed` (equal shards)."""
  world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
  return block.new_empty(block.shape[0] * world, block.shape[1])


def gather_packed(block, out):
  """One collective per control step: every rank's `[n_local, k]` block lands in `out[rank * n_local : ...]` on every
  rank (rank 0 included — the north_star's "gather observations / rewards to rank 0"), in environment order, with no
  temporary buffers and no slice copies (`all_gather_into_tensor`: NCCL ring / NVLS over NVSwitch on GPUs, gloo in the
  CPU tests). Equal shard sizes only; ragged shards go through `gather_to_rank0`."""
  if not dist.is_initialized() or dist.get_world_size() == 1:
    out.copy_(block)
    return out
  if out.shape[0] != block.shape[0] * dist.get_world_size():
    raise ValueError(f'gather_packed needs equal shards: out has {out.shape[0]} rows, block {block.shape[0]} x world {dist.get_world_size()}')
  dist.all_gather_into_tensor(out, block.contiguous())
  return out
