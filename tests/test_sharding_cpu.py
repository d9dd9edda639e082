"""Multi-rank plumbing on CPU: world_size-2/3 gloo groups, the oracle stands in
for the per-rank kernel (the kernel itself is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acrobotics_b200 import sharding


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 31, 32, 33, 1000, 4096, 100_000_007):
        for world in (1, 2, 3, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (lo, hi), (lo2, _) in zip(edges, edges[1:]):
                assert hi == lo2 and (lo % 32 == 0 or lo == n) and lo <= hi
            assert all(hi - lo <= sharding.shard_words(n, world) * 32 for lo, hi in edges)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import acrobotics_b200 as ab
        from acrobotics_b200.synthetic import random_configs, scene16
        from oracle import np_oracle as no

        spec = no.spec_from_robot(ab.Kuka())
        scene = no.scene_arrays(scene16())
        q = torch.from_numpy(random_configs(n, 6, seed=5))

        def check_local(rows):
            hit = no.is_in_collision(spec, scene, rows.numpy())
            pad = (-len(hit)) % 32
            bits = np.concatenate([hit, np.zeros(pad, bool)])
            return torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())

        words = sharding.check_shard(q, check_local)
        np.save(os.path.join(out_dir, f"words{rank}.npy"), words.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1000), (3, 70), (2, 33)])
def test_all_gathered_mask_equals_single_rank(tmp_path, world, n):
    import acrobotics_b200 as ab
    from acrobotics_b200.synthetic import random_configs, scene16
    from oracle import np_oracle as no

    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    spec = no.spec_from_robot(ab.Kuka())
    hit = no.is_in_collision(spec, no.scene_arrays(scene16()), random_configs(n, 6, seed=5))
    for r in range(world):
        w = np.load(tmp_path / f"words{r}.npy").view(np.uint32)
        assert len(w) == (n + 31) // 32
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (bits == hit).all()


# ---------------------------------------------------------------------------
# MaskShards: piece-wise layout with the all-gather of piece k under the kernel of piece k+1
# ---------------------------------------------------------------------------
def test_mask_shards_blocks_tile_the_batch():
    for n in (1, 31, 32, 33, 1000, 4097, 100_000_007):
        for world in (1, 2, 3, 8):
            for pieces in (1, 2, 4):
                shards = [sharding.MaskShards(n, world, r, pieces) for r in range(world)]
                sh = shards[0]
                covered = []
                for k in range(sh.pieces):
                    for r in range(world):
                        lo, hi, w0 = shards[r].block(k)
                        assert w0 == (k * world + r) * sh.per and (lo % 32 == 0 or lo == n)
                        covered.append((lo, hi))
                assert covered[0][0] == 0 and covered[-1][1] == n
                for (lo, hi), (lo2, _) in zip(covered, covered[1:]):
                    assert hi == lo2 and lo <= hi
                assert sh.buffer.numel() == sh.per * world * sh.pieces >= sh.words
    # pieces = 1 is SURVEY 8e's contiguous partition
    for r in range(4):
        assert sharding.MaskShards(1000, 4, r).block(0)[:2] == sharding.shard_range(1000, r, 4)


def _pieces_worker(rank, world, port, n, pieces, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng_bits = np.random.default_rng(77).integers(0, 2, n).astype(bool)  # stand-in verdicts

        def into(lo, hi, out):
            bits = rng_bits[lo:hi]
            bits = np.concatenate([bits, np.zeros((-len(bits)) % 32, bool)])
            w = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())
            out[: w.numel()] = w

        sh = sharding.check_sharded(n, into, pieces=pieces)
        np.save(os.path.join(out_dir, f"p{rank}.npy"), sh.mask().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,pieces", [(2, 5000, 3), (3, 64 * 9 + 5, 2), (2, 40, 4)])
def test_piecewise_gather_equals_single_rank(tmp_path, world, n, pieces):
    mp.spawn(_pieces_worker, args=(world, _free_port(), n, pieces, str(tmp_path)), nprocs=world, join=True)
    want = np.random.default_rng(77).integers(0, 2, n).astype(bool)
    for r in range(world):
        w = np.load(tmp_path / f"p{r}.npy").view(np.uint32)
        assert len(w) == (n + 31) // 32
        assert (np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool) == want).all()


# ---------------------------------------------------------------------------
# planner sweep (C5) sharded over path points; the oracle stands in for the GPU pipeline
# ---------------------------------------------------------------------------
def _oracle_csr(T):
    """CPU stand-in for joint_solutions_csr(..., return_all=True) on (P, S, 4, 4) poses."""
    import acrobotics_b200 as ab
    from acrobotics_b200.synthetic import scene16
    from oracle import np_oracle as no

    spec = no.spec_from_robot(ab.Kuka())
    scene = no.scene_arrays(scene16())
    P, S = T.shape[:2]
    sol, valid = no.ik_kuka(spec, np.asarray(T).reshape(-1, 4, 4))
    hit = np.zeros_like(valid)
    if valid.any():
        hit[valid] = no.is_in_collision(spec, scene, sol[valid])
    free = valid & ~hit
    rows = sol[free]                                        # pose-major, branch-minor order
    counts = free.reshape(P, S * 8).sum(1)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    fbits = np.packbits(free, axis=1, bitorder="little").reshape(P, S)
    vbits = np.packbits(valid, axis=1, bitorder="little").reshape(P, S)
    return (torch.from_numpy(rows.reshape(-1, 6)), torch.from_numpy(off), None,
            torch.from_numpy(vbits), torch.from_numpy(fbits))


def _csr_worker(rank, world, port, P, S, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = torch.from_numpy(np.load(os.path.join(out_dir, "T.npy")))
        res = sharding.joint_solutions_csr_sharded(None, T, None, gather_solutions=True,
                                                   solve_local=lambda Tl: _oracle_csr(Tl.numpy()))
        np.savez(os.path.join(out_dir, f"csr{rank}.npz"), counts=res["counts"].numpy(),
                 free=res["free"].numpy(), q_free=res["q_free"].numpy(),
                 offsets=res["offsets"].numpy(), lo=res["points"][0], hi=res["points"][1],
                 q_local=res["q_local"].numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,P,S", [(2, 7, 5), (3, 5, 4), (2, 1, 3)])
def test_planner_sweep_sharded_over_path_points(tmp_path, world, P, S):
    import acrobotics_b200 as ab
    from acrobotics_b200.synthetic import random_configs
    from oracle import np_oracle as no

    spec = no.spec_from_robot(ab.Kuka())
    T = no.fk(spec, random_configs(P * S, 6, seed=21)).reshape(P, S, 4, 4)
    T[0, 0, :3, 3] *= 5.0  # an unreachable sample
    np.save(tmp_path / "T.npy", T)
    mp.spawn(_csr_worker, args=(world, _free_port(), P, S, str(tmp_path)), nprocs=world, join=True)
    rows, off, _, _, fbits = _oracle_csr(T)
    covered = 0
    for r in range(world):
        g = np.load(tmp_path / f"csr{r}.npz")
        assert np.array_equal(g["counts"], np.diff(off.numpy()))
        assert np.array_equal(g["free"], fbits.numpy())
        assert np.array_equal(g["offsets"], off.numpy())
        assert np.array_equal(g["q_free"], rows.numpy())
        lo, hi = int(g["lo"]), int(g["hi"])
        assert np.array_equal(g["q_local"], rows.numpy()[off[lo]:off[hi]])
        covered += hi - lo
    assert covered == P


# ---------------------------------------------------------------------------
# swept collision (K7) sharded over pairs: a stub robot whose batch call is the numpy oracle
# ---------------------------------------------------------------------------
class _OracleSweepRobot:
    def __init__(self):
        import acrobotics_b200 as ab
        from oracle import np_oracle as no

        self.no, self.spec = no, no.spec_from_robot(ab.Kuka())

    def is_path_in_collision_batch(self, q_start, q_goal, scene, n_samples=0, out=None):
        g = self.no.swept_margins(self.spec, scene, q_start.numpy(), q_goal.numpy(),
                                  n_samples=n_samples or self.no.CCD_SAMPLES)
        hit = ~(g > 0)
        bits = np.concatenate([hit, np.zeros((-len(hit)) % 32, bool)])
        words = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy())
        out[: len(words)] = words


def _swept_inputs(n):
    from acrobotics_b200.synthetic import random_configs

    qs = random_configs(n, 6, seed=11)
    return qs, qs + np.random.default_rng(12).uniform(-0.7, 0.7, qs.shape)


def _swept_worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from acrobotics_b200.synthetic import scene16
        from oracle import np_oracle as no

        qs, qg = _swept_inputs(n)
        words = sharding.is_path_in_collision_sharded(
            _OracleSweepRobot(), torch.from_numpy(qs), torch.from_numpy(qg),
            no.scene_arrays(scene16()), packed=True)
        np.save(os.path.join(out_dir, f"swept{rank}.npy"), words.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 300), (3, 65)])
def test_swept_pairs_sharded_equal_single_rank(tmp_path, world, n):
    import acrobotics_b200 as ab
    from acrobotics_b200.synthetic import scene16
    from oracle import np_oracle as no

    mp.spawn(_swept_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    qs, qg = _swept_inputs(n)
    g = no.swept_margins(no.spec_from_robot(ab.Kuka()), no.scene_arrays(scene16()), qs, qg)
    for r in range(world):
        w = np.load(tmp_path / f"swept{r}.npy").view(np.uint32)
        bits = np.unpackbits(w.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (bits == ~(g > 0)).all()
