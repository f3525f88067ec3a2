"""torch.distributed plumbing for the sharded entry points (one process per GPU, NCCL; gloo in the CPU tests).

The reference has no distributed code; SURVEY.md §8(e) shards the path by rows: encode / decode need no
collective, the bounds need a MIN / MAX all-reduce of two D*nb vectors (update_weights_bounds,
beast/beast_bspline_tokenizer.py:362-378) and the exact quantile needs every rank's coefficient rows
(fit_parameters, :181-220).  A process group is "world" (the default group), a ProcessGroup, or False / None =
local: the tokenizer opts in through set_process_group, it never communicates implicitly.
"""
from typing import List

import torch


WORLD = "world"     # explicit request for the default group where None means "local"


def resolve(process_group=None, implicit=True):
    """(dist, group) when a reduction over more than one rank is due, else (None, None).
    implicit=False: None stays local (per-batch calls must opt in with a group or _dist.WORLD)."""
    if process_group is False or (process_group is None and not implicit):
        return None, None
    if isinstance(process_group, str) and process_group == WORLD:
        process_group = None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return None, None
    if dist.get_world_size(process_group) <= 1:
        return None, None
    return dist, process_group


def allreduce_minmax(lo: torch.Tensor, hi: torch.Tensor, process_group=None, implicit=True):
    """In-place global column min / max over the ranks (order-independent, hence bit-exact)."""
    dist, group = resolve(process_group, implicit)
    if dist is not None:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return lo, hi


def gather_rows(x: torch.Tensor, process_group=None) -> torch.Tensor:
    """Concatenate the ranks' row blocks [n_r, cols] (n_r may differ, 0 allowed) on every rank, rank order."""
    dist, group = resolve(process_group)
    if dist is None:
        return x
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], device=x.device, dtype=torch.int64)
    counts: List[torch.Tensor] = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts_h = [int(c.item()) for c in counts]
    n_max = max(counts_h)
    if n_max == 0:
        return x
    padded = x if x.shape[0] == n_max else torch.cat(
        [x, torch.zeros((n_max - x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts_h)], dim=0)
