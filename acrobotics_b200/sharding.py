"""Multi-GPU sharding of configuration batches.

Every configuration is independent, so a batch of N configurations is cut into
``world`` contiguous row ranges, one per rank (one process per GPU).  The cut
points are multiples of 32 so that a rank's verdict bits are whole ``uint32``
words of the global bit mask; the per-rank word counts are equalised by padding
so that one ``all_gather_into_tensor`` (NCCL over NVLink on the GPU box, gloo in
the CPU tests) assembles the global mask without any repacking.  The robot and
scene descriptors (< 4 KB) are replicated.

Nothing here computes collisions: ``check_shard`` receives the callable that
runs the fused FK -> collision kernel on the local rows.
"""
from typing import Callable, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_words(n: int, world: int) -> int:
    """Mask words (32 configurations each) owned by every rank."""
    words = (int(n) + 31) // 32
    return (words + world - 1) // world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Row range ``[lo, hi)`` of ``rank``; ``lo`` is a multiple of 32 and the
    ranges of all ranks tile ``[0, n)`` exactly (trailing ranks may be empty)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    per = shard_words(n, world) * 32
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def gather_mask(local_words: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather the per-rank mask words into the global mask of ``ceil(n/32)``
    words.  ``local_words`` holds the words of this rank's row range (fewer than
    ``shard_words`` on a trailing rank: it is zero padded)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    total = (int(n) + 31) // 32
    if world == 1:
        return local_words[:total]
    per = shard_words(n, world)
    if local_words.numel() != per:
        padded = torch.zeros(per, dtype=local_words.dtype, device=local_words.device)
        padded[: local_words.numel()] = local_words
        local_words = padded
    out = torch.empty(world * per, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, local_words.contiguous(), group=group)
    return out[:total]


def check_shard(
    q, check_local: Callable[[torch.Tensor], torch.Tensor], group=None
) -> torch.Tensor:
    """Collision mask words of the whole batch ``q`` (N, ndof), replicated on every
    rank: this rank evaluates rows ``shard_range(N, rank, world)`` with
    ``check_local(rows) -> int32 words`` and the words are all-gathered."""
    n = int(q.shape[0])
    if dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    lo, hi = shard_range(n, rank, world)
    rows = q[lo:hi]
    if hi > lo:
        words = check_local(rows)
    else:
        words = torch.zeros(0, dtype=torch.int32, device=getattr(q, "device", "cpu"))
    if not isinstance(words, torch.Tensor):
        words = torch.from_numpy(np.ascontiguousarray(words).view(np.int32))
    return gather_mask(words, n, group)


def is_in_collision_sharded(robot, q, scene, group=None) -> torch.Tensor:
    """``Robot.is_in_collision`` for an (N, ndof) CUDA tensor ``q`` that is
    replicated (or at least addressable) on every rank: bool (N,) on every rank."""
    from . import device

    words = check_shard(
        q, lambda rows: robot.is_in_collision_batch(rows, scene, packed=True), group
    )
    return device.unpack_mask(words, int(q.shape[0]), host=False)
