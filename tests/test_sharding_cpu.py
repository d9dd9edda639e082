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
