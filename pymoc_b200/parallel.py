"""Multi-GPU sharding of an ensemble.

Members never interact, so the path shards trivially: rank ``g`` of ``G`` owns the contiguous
member block ``shard_range(M, g, G)`` and steps it with no collective in the loop.  The only
communication is the gather of diagnostics at the end of a run (NCCL over NVLink on GPUs;
the same code runs on ``gloo`` for the CPU tests).
"""
from __future__ import annotations


def shard_range(M, rank, world):
  """Contiguous block [lo, hi) of rank ``rank``; blocks differ by at most one member."""
  if not (0 <= rank < world):
    raise ValueError('rank %d outside world %d' % (rank, world))
  base, extra = divmod(M, world)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


def gather_members(local, M, group=None):
  """All-gather a per-member tensor ``local`` ([M_local, ...]) into [M, ...] on every rank.

  Blocks may differ by one member (``shard_range``), so the exchange is padded to the
  largest block and trimmed afterwards.
  """
  import torch
  import torch.distributed as dist
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return local
  world = dist.get_world_size(group)
  sizes = [shard_range(M, r, world) for r in range(world)]
  biggest = max(hi - lo for lo, hi in sizes)
  pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
  pad[:local.shape[0]] = local
  out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
  dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
  parts = [out[r * biggest:r * biggest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
  return torch.cat(parts, dim=0)
