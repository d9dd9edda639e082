"""Planner inner loop (pose grid -> IK -> collision filter -> CSR) against the oracle."""
import numpy as np
import pytest
from helpers import spec

import acrobotics_b200 as ab
from acrobotics_b200.path import (NoTolerance, SymmetricTolerance, TolEulerPt, TolPositionPt,
                                  joint_solutions_csr, path_to_joint_solutions)
from acrobotics_b200.synthetic import random_configs, scene16
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu


def _oracle_csr(k, T, scene):
    P, S = T.shape[:2]
    s = spec(k)
    sol, valid = no.ik_kuka(s, T.reshape(-1, 4, 4))
    rows, offsets = [], [0]
    margins = []
    for p in range(P):
        for i in range(p * S, (p + 1) * S):
            for sl in range(8):
                if valid[i, sl]:
                    rows.append(sol[i, sl])
        offsets.append(len(rows))
    rows = np.array(rows).reshape(-1, 6)
    g = no.collision_margins(s, scene, rows) if len(rows) else np.zeros(0)
    return rows, np.array(offsets), g


def test_csr_pipeline_matches_oracle(gpu):
    k = ab.Kuka()
    scene = scene16()
    rng = np.random.default_rng(4)
    P, S = 37, 50
    q = random_configs(P * S, 6, seed=9)
    T = k.fk_batch(q).reshape(P, S, 4, 4)
    T[::6, ::3, :3, 3] *= 4.0  # some unreachable samples
    q_free, off, q_all, valid, free = joint_solutions_csr(k, T, scene, return_all=True)
    assert off.shape == (P + 1,) and off[0] == 0 and off[-1] == len(q_free)
    rows, cand_off, g = _oracle_csr(k, T, scene)
    robust = np.abs(g) > 1e-6
    assert robust.mean() > 0.999
    # expected survivors: candidates with g > 0, in order; compare per path point
    keep = g > 0
    for p in range(P):
        lo, hi = cand_off[p], cand_off[p + 1]
        exp = rows[lo:hi][keep[lo:hi]]
        got = q_free[off[p]:off[p + 1]]
        if robust[lo:hi].all():
            assert got.shape == exp.shape, p
            assert np.abs(got - exp).max(initial=0.0) <= 1e-9
    # the CSR rows are exactly the free-flagged entries of q_all, in order
    bits = np.unpackbits(free.reshape(-1, 1), axis=1, bitorder="little").astype(bool)
    assert np.array_equal(q_all.reshape(-1, 8, 6)[bits], q_free)
    assert not (free & ~valid).any()


def test_csr_large_and_ragged(gpu):
    k = ab.Kuka()
    scene = scene16()
    for P, S in ((1, 1), (3, 33), (2000, 7), (513, 64)):
        q = random_configs(P * S, 6, seed=P)
        T = k.fk_batch(q).reshape(P, S, 4, 4)
        q_free, off, q_all, valid, free = joint_solutions_csr(k, T, scene, return_all=True)
        counts = np.unpackbits(free.reshape(P, S, 1), axis=2, bitorder="little").sum(axis=(1, 2))
        assert np.array_equal(np.diff(off), counts)
        bits = np.unpackbits(free.reshape(-1, 1), axis=1, bitorder="little").astype(bool)
        assert np.array_equal(q_all.reshape(-1, 8, 6)[bits], q_free)
        # every surviving row is collision free and reproduces its pose
        if len(q_free):
            assert not k.is_in_collision_batch(q_free, scene).any()


def test_path_points_to_joint_solutions(gpu):
    """Welding-like line with a rotation tolerance about the tool axis
    (reference tests/test_planning_sampling_based.py style set-up)."""
    k = ab.Kuka()
    scene = ab.Scene([ab.Box(0.5, 0.5, 0.1)], [np.array(
        [[1, 0, 0, 0.8], [0, 1, 0, 0], [0, 0, 1, -0.1], [0, 0, 0, 1.0]])])
    R = np.array([[1.0, 0, 0], [0, -1.0, 0], [0, 0, -1.0]])  # tool pointing down
    path = [TolEulerPt([0.8, y, 0.2], R, [NoTolerance()] * 3,
                       [NoTolerance(), NoTolerance(), SymmetricTolerance(np.pi, 12)])
            for y in np.linspace(-0.2, 0.2, 15)]
    path.append(TolPositionPt([0.8, 0.25, 0.2], R, [SymmetricTolerance(0.05, 3), NoTolerance(),
                                                    NoTolerance()]))
    Q = path_to_joint_solutions(path, k, scene)
    assert len(Q) == len(path)
    for pt, js in zip(path, Q):
        assert js.ndim == 2 and js.shape[1] == 6 and len(js) > 0
        assert not k.is_in_collision_batch(js, scene).any()
        # each solution reproduces one of the sampled poses
        Tq = k.fk_batch(js)
        grid = pt.sample_grid_array()
        err = np.abs(Tq[:, None] - grid[None]).max(axis=(2, 3)).min(axis=1)
        assert err.max() < 1e-6
        # single-point API agrees with the batched one
    one = path[3].to_joint_solutions(k, None, scene)
    assert np.array_equal(one, Q[3])
    with pytest.raises(Exception):
        path_to_joint_solutions([TolPositionPt([5.0, 0, 0], R, [NoTolerance()] * 3)], k, scene)


@pytest.mark.parametrize("cost", ["norm_l1", "norm_l2", "sum_squared", "weighted_sum_squared"])
def test_shortest_path_matches_cpu_dp(gpu, cost):
    from acrobotics_b200.planning import find_shortest_joint_path, shortest_path
    from oracle import dp_oracle

    rng = np.random.default_rng(8)
    sizes = [1, 7, 300, 33, 1000, 2, 64, 513]
    layers = [rng.uniform(-3, 3, (n, 6)) for n in sizes]
    w = np.array([1.0, 2.0, 0.5, 0.1, 3.0, 1.5]) if cost == "weighted_sum_squared" else None
    ref = dp_oracle.shortest_path(layers, cost, w)
    got = shortest_path(layers, cost, w)
    assert got["success"] and abs(got["length"] - ref["length"]) <= 1e-9 * max(1, ref["length"])
    assert np.array_equal(np.array(got["path"]), np.array(ref["path"]))
    # with a state cost per node
    sc = [rng.uniform(0, 2, n) for n in sizes]
    ref = dp_oracle.shortest_path(layers, cost, w, sc)
    got = shortest_path(layers, cost, w, sc)
    assert abs(got["length"] - ref["length"]) <= 1e-9 * max(1, ref["length"])
    assert np.array_equal(np.array(got["path"]), np.array(ref["path"]))
    jp = find_shortest_joint_path(layers, cost, w)
    assert len(jp.joint_positions) == len(sizes) and jp.cost > 0


def test_shortest_path_edge_cases(gpu):
    from acrobotics_b200.planning import find_shortest_joint_path, shortest_path

    one = [np.array([[0.1, 0.2, 0.3]])]
    res = shortest_path(one)
    assert res["success"] and res["length"] == 0.0 and np.array_equal(res["path"][0], one[0][0])
    # ties: identical nodes -> the lowest index is chosen, deterministic
    layers = [np.zeros((5, 2)), np.ones((4, 2)), np.zeros((3, 2))]
    res = shortest_path(layers, "norm_l1")
    assert res["length"] == 4.0
    assert shortest_path([np.zeros((2, 6)), np.zeros((0, 6))])["success"] is False
    with pytest.raises(ValueError):
        find_shortest_joint_path([np.zeros((2, 6)), np.zeros((0, 6))])
    # a straight line in joint space is recovered exactly among distractors
    rng = np.random.default_rng(1)
    line = np.linspace([0, 0, 0, 0, 0, 0], [1, -1, 0.5, 0.2, 0, 2], 40)
    layers = []
    for p in line:
        pts = p + rng.uniform(0.5, 2.0, (50, 6)) * rng.choice([-1, 1], (50, 6))
        k = rng.integers(0, 51)
        layers.append(np.insert(pts, k, p, axis=0))
    res = shortest_path(layers, "sum_squared")
    assert np.allclose(np.array(res["path"]), line)


def test_workspace_envelope(gpu):
    """reference tests/test_workspace_envelope.py:16-23, :54-62 + consistency with the oracle."""
    from acrobotics_b200.workspace_envelope import (EnvelopeSettings, calculate_reachability_batch,
                                                    generate_positions, generate_robot_envelope,
                                                    sample_position)

    pos = np.array([0.1, 0.2, 0.3])
    samples = sample_position(pos, 5, np.random.default_rng(1))
    assert all(np.allclose(tf[:3, 3], pos) for tf in samples)
    assert all(np.any(samples[i] != samples[i - 1]) for i in range(1, 5))
    robot = ab.Kuka()
    settings = EnvelopeSettings(1.0, 10, 8)
    we = generate_robot_envelope(robot, settings, np.random.default_rng(2))
    num_points = int(2 * robot.estimate_max_extension() / settings.sample_distance)
    assert we.shape == (num_points ** 3, 4)
    assert (we[:, 0] >= 0).all() and (we[:, 0] <= 1).all()
    # against the oracle on a finer grid: same poses, IK + self collision on the CPU
    pts = generate_positions(1.2, 5)
    rng = np.random.default_rng(3)
    got = calculate_reachability_batch(robot, pts, 20, 8, rng)
    from acrobotics_b200.workspace_envelope import random_rotations
    rng = np.random.default_rng(3)
    T = np.tile(np.eye(4), (len(pts), 20, 1, 1))
    T[:, :, :3, :3] = random_rotations(len(pts) * 20, rng).reshape(len(pts), 20, 3, 3)
    T[:, :, :3, 3] = pts[:, None, :]
    s = spec(robot)
    sol, valid = no.ik_kuka(s, T.reshape(-1, 4, 4))
    flat = sol[valid]
    g = no.collision_margins(s, None, flat)
    free = np.zeros(valid.shape, bool)
    free[valid] = g > 0
    exp = free.reshape(len(pts), -1).sum(axis=1) / (8 * 20)
    assert np.abs(got - exp).max() <= 1.0 / 160 + 1e-12  # at most one near-contact solution
    assert got.max() > 0.2 and got.min() == 0.0


def test_planner_demo_end_to_end(gpu):
    """examples/planner_demo.py at a small size: the joint path follows the line, is collision
    free and continuous."""
    import importlib.util
    import os

    from conftest import ROOT

    spec_ = importlib.util.spec_from_file_location("planner_demo",
                                                   os.path.join(ROOT, "examples", "planner_demo.py"))
    demo = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(demo)
    q_path, length, (robot, scene, path) = demo.plan(points=60, samples=24, verbose=False)
    assert q_path.shape == (60, 6) and np.isfinite(length)
    assert not robot.is_in_collision_batch(q_path, scene).any()
    T = robot.fk_batch(q_path)
    want = np.array([pt.pos for pt in path])
    assert np.abs(T[:, :3, 3] - want).max() < 1e-9
    # the same graph through the CPU DP: same cost, same joint path
    from acrobotics_b200.path import joint_solutions_csr
    from oracle import dp_oracle

    T_grid = np.stack([pt.sample_grid_array() for pt in path])
    q_free, off = joint_solutions_csr(robot, T_grid, scene)
    ref = dp_oracle.shortest_path([q_free[off[i]:off[i + 1]] for i in range(len(path))], "sum_squared")
    assert abs(ref["length"] - length) <= 1e-9 * max(1.0, length)
    assert np.array_equal(np.array(ref["path"]), q_path)


@pytest.mark.parametrize("cost", ["norm_l1", "norm_l2", "sum_squared", "weighted_sum_squared"])
def test_shortest_path_many_layers_single_launch(gpu, cost):
    """Graphs of >= 32 small layers take the persistent single-launch kernel (grid barrier
    between layers): same path, same length as the CPU DP, ties included."""
    from acrobotics_b200 import _native
    from acrobotics_b200.planning import shortest_path
    from oracle import dp_oracle

    rng = np.random.default_rng(12)
    sizes = list(rng.integers(1, 400, 150)) + [1, 1500, 2, 1920, 1]
    layers = [np.round(rng.uniform(-3, 3, (n, 6)), 1) for n in sizes]  # rounding creates ties
    w = np.array([1.0, 2.0, 0.5, 0.1, 3.0, 1.5]) if cost == "weighted_sum_squared" else None
    sc = [rng.uniform(0, 2, n) for n in sizes]
    for state_cost in (None, sc):
        ref = dp_oracle.shortest_path(layers, cost, w, state_cost)
        before = _native.launch_count()
        got = shortest_path(layers, cost, w, state_cost)
        assert _native.launch_count() - before == 2  # the persistent kernel + the backtrack
        assert got["success"] and abs(got["length"] - ref["length"]) <= 1e-9 * max(1, ref["length"])
        assert np.array_equal(np.array(got["path"]), np.array(ref["path"]))


@pytest.mark.parametrize("ndof", [2, 3, 7])
def test_shortest_path_single_launch_other_widths(gpu, ndof):
    """The persistent kernel's two row-copy paths: one TMA bulk copy per layer when rows and
    base are multiples of 16 bytes (even widths), element-wise cp.async otherwise (odd widths)."""
    from acrobotics_b200 import _native
    from acrobotics_b200.planning import shortest_path
    from oracle import dp_oracle

    rng = np.random.default_rng(30 + ndof)
    sizes = list(rng.integers(1, 300, 40)) + [1900, 3, 700]
    layers = [np.round(rng.uniform(-2, 2, (n, ndof)), 1) for n in sizes]
    for cost in ("sum_squared", "norm_l1"):
        ref = dp_oracle.shortest_path(layers, cost)
        before = _native.launch_count()
        got = shortest_path(layers, cost)
        assert _native.launch_count() - before == 2
        assert abs(got["length"] - ref["length"]) <= 1e-9 * max(1, ref["length"])
        assert np.array_equal(np.array(got["path"]), np.array(ref["path"]))


def test_shortest_path_single_launch_rows_not_16_byte_aligned(gpu):
    """q rows that start 8 bytes off a 16-byte boundary take the element-wise copy path."""
    import torch

    from acrobotics_b200.planning import shortest_path_csr
    from oracle import dp_oracle

    rng = np.random.default_rng(41)
    sizes = list(rng.integers(1, 200, 48))
    layers = [np.round(rng.uniform(-2, 2, (n, 6)), 1) for n in sizes]
    ref = dp_oracle.shortest_path(layers, "sum_squared")
    q = np.concatenate(layers)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    flat = torch.zeros(q.size + 1, dtype=torch.float64, device="cuda")
    flat[1:] = torch.as_tensor(q.ravel()).cuda()
    qd = flat[1:].view(-1, 6)
    assert qd.data_ptr() % 16 == 8
    rows, length = shortest_path_csr(qd, offsets, "sum_squared")
    assert abs(length - ref["length"]) <= 1e-9 * max(1, ref["length"])
    assert np.array_equal(q[rows.cpu().numpy()], np.array(ref["path"]))
