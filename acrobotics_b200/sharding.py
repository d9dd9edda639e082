"""Multi-GPU sharding of the hot path: one process per GPU.

Every configuration (path point, path segment) is independent, so a batch shards with no
data-path exchange; the robot and scene descriptors (< 4 KB) are replicated and the only
collective is the all-gather of the result words (NCCL over NVLink on the GPU box, gloo in the
CPU tests).

Layout (``MaskShards``).  A batch of ``n`` units has ``ceil(n / 32)`` verdict words.  The words
are cut into ``pieces x world`` equal blocks of ``per`` words; block ``(k, r)`` - piece ``k``
of rank ``r`` - covers words ``[(k * world + r) * per, +per)``, i.e. rows
``[(k * world + r) * per * 32, +per * 32)`` of the batch.  With ``pieces = 1`` rank ``r`` owns
the contiguous rows ``[r * N / G, (r + 1) * N / G)`` of SURVEY §8e.  With ``pieces > 1`` the
kernel of piece ``k + 1`` runs while piece ``k`` is all-gathered (the blocks of one piece are
adjacent in the global mask, so every gather is one in-place ``all_gather_into_tensor``: the
kernel writes its words straight into its block of the gather buffer and nothing is repacked).
Measured on eight B200s: the fused kernel is persistent and fills every SM, so
the NCCL kernel of piece ``k`` waits for the compute kernel of piece ``k + 1`` instead of
overlapping it - ``pieces = 2`` is 4 % slower than ``pieces = 1``, which is therefore the
default; the pieces layout stays for callers whose local kernel leaves SMs free.

Nothing here computes collisions: the callers pass the callable that runs the fused
FK -> collision kernel on the local rows.
"""
from typing import Callable, List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def _rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_words(n: int, world: int) -> int:
    """Mask words (32 configurations each) owned by every rank (``pieces = 1``)."""
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


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced range of ``n`` units (path points, segments) for ``rank``."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (n * rank) // world, (n * (rank + 1)) // world


class MaskShards:
    """Partition of the verdict words of ``n`` units over ``world`` ranks in ``pieces`` pieces,
    plus the gather buffer every rank's kernel writes into (see the module docstring)."""

    def __init__(self, n: int, world: int, rank: int, pieces: int = 1, device="cpu",
                 buffer: torch.Tensor = None):
        if not (0 <= rank < world):
            raise ValueError(f"rank {rank} outside world of {world}")
        self.n, self.world, self.rank = int(n), int(world), int(rank)
        self.words = (self.n + 31) // 32
        self.pieces = max(1, min(int(pieces), max(1, self.words // max(1, world))))
        self.per = max(1, -(-self.words // (self.world * self.pieces)))
        total = self.per * self.world * self.pieces
        if buffer is None:
            buffer = torch.zeros(total, dtype=torch.int32, device=device)
        elif buffer.numel() < total or buffer.dtype != torch.int32:
            raise ValueError(f"gather buffer needs {total} int32 words")
        self.buffer = buffer[:total]

    def block(self, piece: int, rank: int = None) -> Tuple[int, int, int]:
        """(row_lo, row_hi, word_lo) of block ``(piece, rank)``; rows may be empty at the tail."""
        r = self.rank if rank is None else rank
        w0 = (piece * self.world + r) * self.per
        lo = min(self.n, w0 * 32)
        hi = min(self.n, (w0 + self.per) * 32)
        return lo, hi, w0

    def local_words(self, piece: int) -> torch.Tensor:
        """This rank's block of piece ``piece`` inside the gather buffer (a view)."""
        w0 = (piece * self.world + self.rank) * self.per
        return self.buffer[w0:w0 + self.per]

    def piece_words(self, piece: int) -> torch.Tensor:
        w0 = piece * self.world * self.per
        return self.buffer[w0:w0 + self.world * self.per]

    def gather_piece(self, piece: int, group=None, async_op=False):
        """In-place all-gather of one piece (no-op on a single rank)."""
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(self.piece_words(piece), self.local_words(piece),
                                           group=group, async_op=async_op)

    def mask(self) -> torch.Tensor:
        """The global mask words (valid after every piece was gathered)."""
        return self.buffer[: self.words]


def check_sharded(n: int, check_rows_into: Callable[[int, int, torch.Tensor], None], group=None,
                  pieces: int = 1, device="cpu", shards: MaskShards = None) -> MaskShards:
    """Evaluate ``n`` units across the ranks of ``group``.

    ``check_rows_into(lo, hi, out_words)`` must evaluate units ``[lo, hi)`` (``lo`` a multiple
    of 32) and write their verdict words into ``out_words`` (``ceil((hi - lo) / 32)`` leading
    entries; the rest of the block stays as it is - zero on a fresh buffer).  The call is
    asynchronous on the current CUDA stream; each piece is all-gathered while the next one is
    computed.  Returns the ``MaskShards`` whose ``mask()`` holds all verdicts on every rank."""
    rank, world = _rank_world(group)
    sh = shards if shards is not None else MaskShards(n, world, rank, pieces, device)
    pending = []
    for k in range(sh.pieces):
        lo, hi, _ = sh.block(k)
        if hi > lo:
            check_rows_into(lo, hi, sh.local_words(k))
        work = sh.gather_piece(k, group, async_op=sh.pieces > 1)
        if work is not None:
            pending.append(work)
    for work in pending:
        work.wait()  # the current stream waits for the collective's stream; no host sync on NCCL
    return sh


# ---------------------------------------------------------------------------
# generic (allocating) helpers kept for callers that return their words
# ---------------------------------------------------------------------------
def gather_mask(local_words: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather the per-rank mask words into the global mask of ``ceil(n/32)``
    words.  ``local_words`` holds the words of this rank's row range (fewer than
    ``shard_words`` on a trailing rank: it is zero padded)."""
    rank, world = _rank_world(group)
    total = (int(n) + 31) // 32
    if world == 1:
        return local_words[:total]
    sh = MaskShards(n, world, rank, 1, local_words.device)
    sh.local_words(0)[: local_words.numel()] = local_words
    sh.gather_piece(0, group)
    return sh.mask()


def check_shard(q, check_local: Callable[[torch.Tensor], torch.Tensor], group=None) -> torch.Tensor:
    """Collision mask words of the whole batch ``q`` (N, ndof), replicated on every
    rank: this rank evaluates rows ``shard_range(N, rank, world)`` with
    ``check_local(rows) -> int32 words`` and the words are all-gathered."""
    n = int(q.shape[0])

    def into(lo, hi, out):
        words = check_local(q[lo:hi])
        if not isinstance(words, torch.Tensor):
            words = torch.from_numpy(np.ascontiguousarray(words).view(np.int32))
        out[: words.numel()] = words.to(out.device)

    return check_sharded(n, into, group, 1, getattr(q, "device", "cpu")).mask()


# ---------------------------------------------------------------------------
# the product's multi-GPU entry points
# ---------------------------------------------------------------------------
def is_in_collision_sharded(robot, q, scene, group=None, pieces: int = 1, packed: bool = False,
                            shards: MaskShards = None):
    """``Robot.is_in_collision`` for an (N, ndof) CUDA tensor ``q`` that every rank holds
    (strong scaling: N is the whole job): bool (N,) - or, with ``packed``, the mask words - on
    every rank.  The kernel writes straight into the gather buffer."""
    from . import device

    n = int(q.shape[0])

    def into(lo, hi, out):
        robot.is_in_collision_batch(q[lo:hi], scene, packed=True, out=out)

    sh = check_sharded(n, into, group, pieces, q.device, shards)
    return sh.mask() if packed else device.unpack_mask(sh.mask(), n, host=False)


def is_in_collision_local_shard(robot, q_local, scene, group=None, pieces: int = 1,
                                shards: MaskShards = None) -> MaskShards:
    """Weak-scaling form: every rank brings its OWN (n, ndof) rows; the job is the
    ``world * n`` configurations whose row order is the block order of ``MaskShards`` (local row
    ``i`` of rank ``r`` is global row ``((i // B) * world + r) * B + i % B``, ``B = per * 32``).
    Requires ``n`` to be a multiple of ``32 * pieces``.  Returns the shards (``mask()`` = all
    ``world * n`` verdicts on every rank)."""
    rank, world = _rank_world(group)
    n = int(q_local.shape[0])
    sh = shards if shards is not None else MaskShards(n * world, world, rank, pieces, q_local.device)
    if sh.per * 32 * sh.pieces != n:
        raise ValueError(f"local batch of {n} rows does not split into {sh.pieces} pieces of whole words")
    rows = sh.per * 32
    pending = []
    for k in range(sh.pieces):
        robot.is_in_collision_batch(q_local[k * rows:(k + 1) * rows], scene, packed=True,
                                    out=sh.local_words(k))
        work = sh.gather_piece(k, group, async_op=sh.pieces > 1)
        if work is not None:
            pending.append(work)
    for work in pending:
        work.wait()
    return sh


def is_path_in_collision_discrete_sharded(robot, q_start, q_goal, scene, max_q_step=0.1,
                                          group=None, packed: bool = False):
    """``Robot.is_path_in_collision_discrete`` for (E, ndof) segment arrays every rank holds:
    segments are sharded like configurations, the verdict words all-gathered."""
    from . import device

    e = int(q_start.shape[0])

    def into(lo, hi, out):
        robot.is_path_in_collision_discrete_batch(q_start[lo:hi], q_goal[lo:hi], scene, max_q_step,
                                                  packed=True, out=out)

    sh = check_sharded(e, into, group, 1, q_start.device)
    return sh.mask() if packed else device.unpack_mask(sh.mask(), e, host=False)


def is_path_in_collision_sharded(robot, q_start, q_goal, scene, n_samples=0, group=None,
                                 packed: bool = False):
    """Swept collision ``Robot.is_path_in_collision`` (robot.py:285-317) for (E, ndof) pair
    arrays every rank holds: pairs are sharded like configurations (rank r sweeps its
    contiguous, word-aligned block straight into its slice of the gather buffer), the verdict
    words all-gathered.  No other exchange: the pairs are independent."""
    from . import device

    e = int(q_start.shape[0])

    def into(lo, hi, out):
        robot.is_path_in_collision_batch(q_start[lo:hi], q_goal[lo:hi], scene, n_samples=n_samples,
                                         out=out)

    sh = check_sharded(e, into, group, 1, q_start.device)
    return sh.mask() if packed else device.unpack_mask(sh.mask(), e, host=False)


def joint_solutions_csr_sharded(robot, T, scene, group=None, gather_solutions: bool = False,
                                solve_local: Callable = None, trim: bool = True):
    """Planner sweep (``path_to_joint_solutions``, reference planning/sampling_based.py:94-106)
    over the ranks: ``T`` (P, S, 4, 4) is held by every rank, rank ``r`` evaluates the path
    points ``split_range(P, r, world)`` (IK -> collision filter -> CSR compaction on its GPU)
    and the per-point solution counts and the per-pose free-branch masks are all-gathered.

    Returns a dict: ``counts`` (P,) int64 and ``free`` (P, S) uint8 on every rank;
    ``q_local`` / ``offsets_local`` / ``points`` = this rank's CSR block and its point range;
    with ``gather_solutions`` also ``q_free`` / ``offsets`` = the whole CSR on every rank
    (padded all-gather of the rows, then one compaction).  ``trim=False``: ``q_local`` keeps its
    full capacity (rows beyond ``offsets_local[-1]`` undefined) and nothing is read back to the
    host, so the call is asynchronous."""
    from .path.path_pt import joint_solutions_csr

    rank, world = _rank_world(group)
    P, S = int(T.shape[0]), int(T.shape[1])
    lo, hi = split_range(P, rank, world)
    per = -(-P // world) if world > 1 else P
    dev = T.device if isinstance(T, torch.Tensor) else "cpu"
    solve = solve_local or (lambda Tl: joint_solutions_csr(robot, Tl, scene, return_all=True,
                                                           trim=trim or gather_solutions))
    if hi > lo:
        q_local, off_local, _, _, free_local = solve(T[lo:hi])
    else:
        q_local = torch.zeros((0, 6), dtype=torch.float64, device=dev)
        off_local = torch.zeros(1, dtype=torch.int64, device=dev)
        free_local = torch.zeros((0, S), dtype=torch.uint8, device=dev)
    q_local, off_local, free_local = (torch.as_tensor(x).to(dev) for x in (q_local, off_local, free_local))
    counts_local = off_local[1:] - off_local[:-1]
    out = {"q_local": q_local, "offsets_local": off_local, "points": (lo, hi)}
    if world == 1:
        out.update(counts=counts_local, free=free_local)
        if gather_solutions:
            out.update(q_free=q_local, offsets=off_local)
        return out
    # ranges from split_range differ by at most one point: pad every rank to `per` points
    cbuf = torch.zeros(world * per, dtype=torch.int64, device=dev)
    fbuf = torch.zeros((world * per, S), dtype=torch.uint8, device=dev)
    cbuf[rank * per:rank * per + (hi - lo)] = counts_local
    fbuf[rank * per:rank * per + (hi - lo)] = free_local
    dist.all_gather_into_tensor(cbuf, cbuf[rank * per:(rank + 1) * per].clone(), group=group)
    dist.all_gather_into_tensor(fbuf, fbuf[rank * per:(rank + 1) * per].clone(), group=group)
    keep = _unpad_index(P, world, per, dev)
    counts = cbuf[keep]
    out.update(counts=counts, free=fbuf[keep])
    if gather_solutions:
        ndof = int(q_local.shape[1])
        totals = cbuf.view(world, per).sum(dim=1)            # rows per rank
        cap = int(totals.max().item())
        rows = torch.zeros((world, max(cap, 1), ndof), dtype=q_local.dtype, device=dev)
        mine = torch.zeros((max(cap, 1), ndof), dtype=q_local.dtype, device=dev)
        mine[: q_local.shape[0]] = q_local
        dist.all_gather_into_tensor(rows.view(-1), mine.view(-1), group=group)
        tl = totals.tolist()
        out["q_free"] = torch.cat([rows[r, : tl[r]] for r in range(world)])
        off = torch.zeros(P + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(counts, 0)
        out["offsets"] = off
    return out


def _unpad_index(P: int, world: int, per: int, dev) -> torch.Tensor:
    """Indices into the (world * per)-long padded gather that enumerate the P points in order."""
    idx: List[torch.Tensor] = []
    for r in range(world):
        lo, hi = split_range(P, r, world)
        idx.append(torch.arange(r * per, r * per + (hi - lo), device=dev))
    return torch.cat(idx)
