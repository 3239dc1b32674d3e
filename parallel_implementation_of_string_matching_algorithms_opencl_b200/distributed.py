"""Multi-GPU sharding of the scan: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-device path (one OpenCL device, BoyreMoore.cpp:217-219); its only
partitioning is the host's word split into P ranges WITHOUT overlap, which loses matches that
straddle a seam (BoyreMoore.cpp:119-141).  Here rank r owns the match START
positions [lo_r, hi_r) and reads (m-1) bytes of halo behind hi_r, so every occurrence is reported
exactly once by the rank that owns its start, with its global offset (pos_base = lo_r).

Exchange step (the only collective on the path): all_reduce(sum) of the hit counts, all_gather of
the per-rank counts, then the per-rank position lists are sent to rank 0, whose rank-order
concatenation is already globally ascending.  Over NCCL this runs on NVLink/NVSwitch; the same
code runs over gloo on CPU tensors in the tests.
"""
from __future__ import annotations

from typing import Callable, Optional

SHARD_ALIGN = 16  # shard starts are 16-byte aligned so each rank's TMA tiles stay aligned


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Start positions [lo, hi) owned by `rank` (contiguous, disjoint, covering [0, n_total))."""
    per = -(-n_total // world)
    per = -(-per // SHARD_ALIGN) * SHARD_ALIGN
    lo = min(n_total, rank * per)
    hi = min(n_total, (rank + 1) * per)
    return lo, hi


def shard_read_range(n_total: int, m: int, lo: int, hi: int) -> tuple[int, int]:
    """Bytes [lo, end) a rank must hold: its own range plus the (m-1)-byte halo."""
    return lo, min(n_total, hi + max(m - 1, 0))


def combine_hits(count_local: int, positions_local, *, group=None, device=None, dst: int = 0):
    """Exchange step.  Returns (total_count, per_rank_counts, positions_on_dst_or_None).

    positions_local: 1-D int64 tensor of this rank's GLOBAL positions (ascending), on `device`.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = device if device is not None else positions_local.device
    held_local = 0 if positions_local is None else int(positions_local.numel())  # may be capped below the count
    mine = torch.tensor([int(count_local), held_local], dtype=torch.int64, device=device)

    total = mine[:1].clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)          # every rank learns the global count
    both = torch.empty(2 * world, dtype=torch.int64, device=device)
    _all_gather_list(both, mine, group)                                  # per-rank counts and list lengths
    both_host = both.view(world, 2).tolist()
    counts_host = [int(c) for c, _ in both_host]
    held_host = [int(h) for _, h in both_host]

    gathered = None
    if rank == dst:
        gathered = torch.empty(sum(held_host), dtype=torch.int64, device=device)
        off, reqs = 0, []
        for r in range(world):
            k = held_host[r]
            if k:
                if r == dst:
                    gathered[off: off + k].copy_(positions_local[:k])
                else:
                    reqs.append(dist.irecv(gathered[off: off + k], src=r, group=group))
            off += k
        for q in reqs:
            q.wait()
    elif held_local:
        dist.send(positions_local[:held_local].contiguous(), dst=dst, group=group)
    return int(total.item()), counts_host, gathered


def _all_gather_list(out, mine, group):
    import torch
    import torch.distributed as dist

    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, mine, group=group)
    out.copy_(torch.cat(parts))


def sharded_search(shard_text, lo: int, pattern: bytes, *, max_positions: int, group=None, dst: int = 0,
                   variant="auto", local_scan: Optional[Callable] = None):
    """Scan this rank's shard and run the exchange step.

    shard_text holds the bytes [lo, end) of the global text (own range + halo, see
    shard_read_range).  Returns (total_count, per_rank_counts, positions_on_dst_or_None,
    local_stats).  `local_scan(shard_text, pattern, pos_base, max_positions)` ->
    (count, positions_tensor, stats) defaults to the CUDA scan; the CPU tests inject a checker.
    """
    if local_scan is None:
        from .host import search_device

        def local_scan(text, pat, pos_base, cap):
            return search_device(text, pat, max_positions=cap, pos_base=pos_base, variant=variant)

    count, pos, stats = local_scan(shard_text, pattern, lo, max_positions)
    total, counts, gathered = combine_hits(count, pos, group=group, device=shard_text.device, dst=dst)
    return total, counts, gathered, stats
