"""Work split of a batch of independent frames across the GPUs of one box.

A batch of frames shards naturally: frame i is coded by exactly one rank, no sample or stream byte ever moves between
GPUs.  The only exchange is the control plane -- every rank learns every frame's compressed size so that a global
offset table (where would frame i start in one concatenated output?) can be built."""
from __future__ import annotations


def frame_range(total_frames: int, world_size: int, rank: int) -> range:
    """Contiguous block of frames owned by `rank`; the first (total % world) ranks get one extra frame."""
    base, extra = divmod(total_frames, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_sizes(local_sizes, total_frames: int, dist=None, device=None):
    """All-gathers per-frame stream sizes (uneven shards allowed). Returns the list of all sizes in frame order."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_sizes)
    world = dist.get_world_size()
    longest = max(len(frame_range(total_frames, world, r)) for r in range(world))
    padded = torch.zeros(longest, dtype=torch.int64, device=device)
    padded[: len(local_sizes)] = torch.tensor(list(local_sizes), dtype=torch.int64, device=device)
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    out = []
    for r in range(world):
        out += gathered[r][: len(frame_range(total_frames, world, r))].tolist()
    return out


def offset_table(sizes):
    """Exclusive prefix sum: byte offset of every frame in the concatenated output, plus the total."""
    offsets, total = [], 0
    for s in sizes:
        offsets.append(total)
        total += int(s)
    return offsets, total
