"""The multi-GPU entry points on one GPU (world size 1): the in-place, piece-wise path the
benchmark times must give exactly the verdicts of the plain batched call.  The N > 1 plumbing
(block layout, piece-wise all-gather, planner sweep over path points) is covered on CPU with
gloo groups in tests/test_sharding_cpu.py and on NCCL by bench.py under torchrun."""
import numpy as np
import pytest
import torch

import acrobotics_b200 as ab
from acrobotics_b200 import sharding
from acrobotics_b200.path import joint_solutions_csr
from acrobotics_b200.synthetic import random_configs_device, scene16

pytestmark = pytest.mark.gpu


def test_in_place_pieces_equal_plain_call(gpu):
    k, scene = ab.Kuka(), scene16()
    n = 32 * 4 * 1000
    q = random_configs_device(n, 6, seed=3)
    want = k.is_in_collision_batch(q, scene, packed=True)
    for pieces in (1, 2, 4):
        sh = sharding.is_in_collision_local_shard(k, q, scene, pieces=pieces)
        assert sh.pieces == pieces and torch.equal(sh.mask(), want)
        # the gather buffer is reusable: a second pass overwrites, it does not accumulate
        sh.buffer.fill_(-1)
        sharding.is_in_collision_local_shard(k, q, scene, pieces=pieces, shards=sh)
        assert torch.equal(sh.mask(), want)
    with pytest.raises(ValueError):
        sharding.is_in_collision_local_shard(k, q[: n - 32], scene, pieces=4)
    for m in (n, n - 5, 33, 1):
        got = sharding.is_in_collision_sharded(k, q[:m], scene, pieces=3, packed=True)
        assert torch.equal(got, k.is_in_collision_batch(q[:m], scene, packed=True))
    assert torch.equal(sharding.is_in_collision_sharded(k, q[:1000], scene),
                       k.is_in_collision_batch(q[:1000], scene))


def test_out_argument_is_validated(gpu):
    k, scene = ab.Kuka(), scene16()
    q = random_configs_device(100, 6, seed=4)
    with pytest.raises(ValueError):
        k.is_in_collision_batch(q, scene, out=torch.zeros(2, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        k.is_in_collision_batch(q, scene, out=torch.zeros(4, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        k.is_in_collision_batch(q, scene, out=torch.zeros(4, dtype=torch.int32))
    out = torch.full((8,), -1, dtype=torch.int32, device="cuda")
    res = k.is_in_collision_batch(q, scene, out=out)
    assert res.data_ptr() == out.data_ptr()
    assert torch.equal(out[:4], k.is_in_collision_batch(q, scene, packed=True))
    assert (out[4:] == -1).all()  # nothing beyond ceil(N / 32) words is touched


def test_path_segments_and_planner_sweep_sharded(gpu):
    k, scene = ab.Kuka(), scene16()
    qs = random_configs_device(5000, 6, seed=5)
    qg = qs + 0.3 * torch.randn_like(qs)
    want = k.is_path_in_collision_discrete_batch(qs, qg, scene, 0.1)
    got = sharding.is_path_in_collision_discrete_sharded(k, qs, qg, scene, 0.1)
    assert torch.equal(got, want)
    want = k.is_path_in_collision_batch(qs, qg, scene)
    assert torch.equal(sharding.is_path_in_collision_sharded(k, qs, qg, scene), want)
    assert 0.05 < float(want.float().mean()) < 0.95
    P, S = 40, 25
    T = k.fk_batch(random_configs_device(P * S, 6, seed=6)).view(P, S, 4, 4)
    q_free, off, _, _, free = joint_solutions_csr(k, T, scene, return_all=True)
    res = sharding.joint_solutions_csr_sharded(k, T, scene, gather_solutions=True)
    assert torch.equal(res["offsets"], off) and torch.equal(res["q_free"], q_free)
    assert torch.equal(res["free"], free) and torch.equal(res["counts"], off[1:] - off[:-1])
    assert res["points"] == (0, P)
