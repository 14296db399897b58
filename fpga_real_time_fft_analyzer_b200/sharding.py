"""Channel sharding across the GPUs of one box.

Channels are independent end to end, so rank r owns a contiguous range and the hot
path has no inter-GPU traffic: the window ROM (32 KiB) and the two coefficient banks
(24 B) are replicated.  The only collective is optional - gathering the packed
spectra onto one rank when a single consumer must own the display - and uses
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def channel_range(total_channels: int, rank: int, world: int):
    """Contiguous, balanced split: the first (total % world) ranks get one extra."""
    if not (0 <= rank < world) or total_channels < 0:
        raise ValueError("bad rank/world")
    base, extra = divmod(total_channels, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_frames(local_frames, total_channels: int, dst: int = 0, group=None):
    """Gather each rank's packed frames [c_local, 4N] uint8 onto rank `dst` in channel
    order.  Returns the [total_channels, 4N] tensor on dst, None elsewhere.
    Ranges may be ragged, so ranks pad to the largest shard for all_gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [channel_range(total_channels, r, world) for r in range(world)]
    cmax = max(b - a for a, b in sizes)
    width = local_frames.shape[1]
    padded = torch.zeros((cmax, width), dtype=local_frames.dtype, device=local_frames.device)
    padded[: local_frames.shape[0]] = local_frames
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
