"""Multi-GPU sharding of candidate sites.

Candidates are independent end to end, so the path shards with NO data-path collective: rank r owns
the contiguous candidate range ``shard_bounds(n, world, r)`` and runs encoder + AFF + NEG + posterior
on it.  The only exchange is ONE gather of the per-candidate results (probabilities / posteriors,
64-96 B per candidate) to rank 0, which writes the chunk's predict / VCF rows in candidate order.
Backend: NCCL over NVLink on the B200 box, gloo in the CPU tests.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced ranges: the first ``n % world`` ranks get one extra candidate."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world: int):
    return [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, group=None):
    """Gather per-candidate rows [n_local, ...] from every rank into [n_total, ...] on ``dst``
    (rank order == candidate order).  Shards are padded to the largest shard so a single
    ``all_gather_into_tensor`` (one NCCL collective) moves everything.  Returns None off ``dst``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    width = max(sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if rank != dst:
        return None
    out = out.reshape((world, width) + tuple(local.shape[1:]))
    return torch.cat([out[r, :sizes[r]] for r in range(world)], dim=0)


def exchange_sizes(n_local: int, device, group=None):
    """Shard lengths of every rank (one tiny all_gather + a host read: do it once per batch, outside any hot loop)."""
    world = dist.get_world_size(group)
    n = torch.tensor([n_local], dtype=torch.int64, device=device)
    sizes = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    return [int(v) for v in sizes.tolist()]


def gather_rows_padded(local: torch.Tensor, sizes=None, dst: int = 0, group=None):
    """The same single collective for shards of arbitrary lengths ``sizes`` (from ``exchange_sizes``; exchanged here when
    omitted): rows are padded to the longest shard.  Returns the concatenated rows on ``dst``, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if sizes is None:
        sizes = exchange_sizes(local.shape[0], local.device, group)
    width = max(sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if rank != dst:
        return None
    out = out.reshape((world, width) + tuple(local.shape[1:]))
    return torch.cat([out[r, :sizes[r]] for r in range(world)], dim=0)
