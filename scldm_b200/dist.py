"""Multi-GPU generation: cells are independent, so the batch is split contiguously over ranks (one process
per GPU), every rank runs the full path on its slice, and the only collective is the final output gather
(SURVEY.md §8e).  RNG streams are keyed by the *global* cell index, so the generated cells do not depend on
the number of GPUs.  The reference itself is single-GPU for sampling (`inference.py:68-70`)."""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra cell."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all-gather row blocks of unequal length (shard_range layout) into the full [n_total, ...] tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    max_rows = max(b - a for a, b in sizes)
    pad = torch.zeros((max_rows, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def sample_sharded(ldm, condition: dict | None, guidance_weight: dict | None, batch_size: int, genes: torch.Tensor,
                   gather: bool = True, **kw):
    """`LatentDiffusion.sample` for a global batch split over the ranks of the default process group.

    Every rank passes the *global* `condition` / `genes`; it generates cells [start, stop) of the batch.  With
    `gather=True` the unconditional and guided halves are all-gathered so that every rank returns the same
    (counts (2B,G), z (2B,M,L)) the single-GPU call would; otherwise the local (2b,G) block is returned."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    a, b = shard_range(batch_size, rank, world)
    cond = None if condition is None else {k: v[a:b] for k, v in condition.items()}
    base = ldm.cells_generated
    out = ldm.sample(cond, guidance_weight, b - a, genes[a:b], cell_offset=base + a, **kw)
    ldm.cells_generated = base + batch_size
    if not gather or world == 1:
        return out
    n = b - a
    full = []
    for tsr in out:  # rows [0,n) unconditional, [n,2n) guided on every rank
        full.append(torch.cat([gather_rows(tsr[:n].contiguous(), batch_size), gather_rows(tsr[n:].contiguous(), batch_size)], dim=0))
    return tuple(full)
