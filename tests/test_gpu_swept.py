"""K7 swept collision (``acro_swept_cc_f64``; reference ``Robot.is_path_in_collision``,
robot.py:285-317) through the product API: verbatim-reference golden vectors, the numpy
oracle with margin bookkeeping, and size-independent properties at 1 M pairs."""
import numpy as np
import pytest
from conftest import golden
from helpers import product_scene, spec

import acrobotics_b200 as ab
from acrobotics_b200.synthetic import random_configs, scene16
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu

MARGIN = 1e-6  # north_star: verdicts bit-exact outside a 1e-6 m margin, the rest reported


def _kuka(g, tool=True, base=False):
    k = ab.Kuka()
    if tool:
        k.tool = ab.Tool([ab.Box(*s) for s in g["tool_sides"]], list(g["tool_tfs"]), g["tool_tip"])
    if base:
        k.tf_base = g["full_base_tf"]
        k.geometry_base = product_scene(g["full_base_sides"], g["full_base_tfs"])
    return k


def test_reference_known_answers(gpu):
    """tests/test_continuous_cc.py:35-83: the stateless verdicts (True, True, False; the
    reference's third True comes from python-fcl's sticky result object, see
    tests/test_swept_cpu.py)."""
    g = golden("ccd.npz")
    k = _kuka(g)
    got = [
        k.is_path_in_collision(g["known_q_start"][i], g["known_q_goal"][i],
                               product_scene(g["known_sides"][i], g["known_tfs"][i]))
        for i in range(3)
    ]
    assert got == g["known_stateless"].tolist() == [True, True, False]


@pytest.mark.parametrize("case", ["s3", "s16", "full", "rail"])
def test_verbatim_reference_goldens(gpu, case):
    g = golden("ccd.npz")
    k = ab.KukaOnRail() if case == "rail" else _kuka(g, base=case == "full")
    scene = product_scene(g["s3_sides"], g["s3_tfs"]) if case in ("s3", "full") else scene16()
    hit, near = k.is_path_in_collision_batch(g[case + "_q_start"], g[case + "_q_goal"], scene,
                                             return_near=True, margin=MARGIN)
    ref = g[case + "_hit"]
    assert np.array_equal(hit[~near], ref[~near])
    assert near.sum() <= 2
    # without the margin bookkeeping the same kernel instantiation is not used: check it too
    plain = k.is_path_in_collision_batch(g[case + "_q_start"], g[case + "_q_goal"], scene)
    assert np.array_equal(plain[~near], ref[~near])


@pytest.mark.parametrize("tool", [False, True])
def test_against_numpy_oracle(gpu, tool):
    g = golden("ccd.npz")
    k = _kuka(g, tool=tool)
    scene = scene16()
    n = 20000 if not tool else 8000
    qs = random_configs(n, 6, seed=31 + tool)
    rng = np.random.default_rng(5 + tool)
    qg = qs + rng.uniform(-1.0, 1.0, qs.shape) * rng.choice([0.05, 0.5, 2.5], size=(n, 1))
    hit, near = k.is_path_in_collision_batch(qs, qg, scene, return_near=True, margin=MARGIN)
    gm = no.swept_margins(spec(k), scene, qs, qg)
    o_hit, o_near = ~(gm > 0), np.abs(gm) <= MARGIN
    bucket = near | o_near
    mismatches = int((hit != o_hit)[~bucket].sum())
    print(f"swept parity: {n} pairs, hit rate {o_hit.mean():.3f}, mismatches outside the "
          f"margin {mismatches}, margin bucket {int(bucket.sum())}")
    assert mismatches == 0
    assert bucket.sum() <= max(2, n // 5000)
    assert 0.05 < o_hit.mean() < 0.95


def test_two_samples_are_the_end_poses(gpu):
    """n_samples = 2 tests exactly the start and the goal pose: the OR of two discrete
    checks without self collision (robot.py:288-289)."""
    k = ab.Kuka()
    k.do_check_self_collision = False
    scene = scene16()
    n = 200_000
    qs = random_configs(n, 6, seed=77)
    qg = random_configs(n, 6, seed=78)
    swept, near = k.is_path_in_collision_batch(qs, qg, scene, n_samples=2, return_near=True,
                                               margin=MARGIN)
    a, na = k.is_in_collision_batch(qs, scene, return_near=True, margin=MARGIN)
    b, nb = k.is_in_collision_batch(qg, scene, return_near=True, margin=MARGIN)
    ok = ~(near | na | nb)
    assert np.array_equal(swept[ok], (a | b)[ok])
    assert ok.mean() > 0.9999


def test_no_motion_equals_discrete_check_1m(gpu):
    """Size-independent property at 1 M pairs: q_goal = q_start sweeps nothing, every sample
    is the start pose."""
    import torch

    k = ab.Kuka()
    k.do_check_self_collision = False
    scene = scene16()
    n = 1_000_000
    q = torch.as_tensor(random_configs(n, 6, seed=123)).cuda()
    swept, near = k.is_path_in_collision_batch(q, q.clone(), scene, return_near=True, margin=MARGIN)
    disc, nd = k.is_in_collision_batch(q, scene, return_near=True, margin=MARGIN)
    ok = ~(near | nd)
    assert bool((swept[ok] == disc[ok]).all())
    assert float(ok.float().mean()) > 0.9999
    # a motion can only add collisions to those of its end poses (both are samples)
    q2 = torch.as_tensor(random_configs(n, 6, seed=124)).cuda()
    moved, nm = k.is_path_in_collision_batch(q, q2, scene, return_near=True, margin=MARGIN)
    assert bool((moved | ~disc | nm | nd).all())


def test_edge_cases(gpu):
    k = ab.Kuka()
    scene = scene16()
    assert k.is_path_in_collision_batch(np.zeros((0, 6)), np.zeros((0, 6)), scene).shape == (0,)
    # empty scene / no scene: nothing to hit (self collision is not part of the sweep)
    q = random_configs(100, 6, seed=1)
    assert not k.is_path_in_collision_batch(q, q[::-1].copy(), ab.Scene([], [])).any()
    with pytest.raises(ValueError):
        k.is_path_in_collision_batch(q, q[:50], scene)
    with pytest.raises(Exception):
        k.is_path_in_collision_batch(q, q, scene, n_samples=1)
    # ragged tail (not a multiple of the warp / CTA size) and an odd number of joints
    arm = ab.AnthropomorphicArm()
    q3 = random_configs(1001, 3, seed=2)
    q4 = q3 + 0.4
    hit, near = arm.is_path_in_collision_batch(q3, q4, scene, return_near=True, margin=MARGIN)
    gm = no.swept_margins(spec(arm), scene, q3, q4)
    ok = ~(near | (np.abs(gm) <= MARGIN))
    assert np.array_equal(hit[ok], ~(gm > 0)[ok])


def test_persistent_and_per_thread_kernels_agree(gpu):
    """The optional persistent kernel (lanes take pairs from a counter, one box set-up or one
    sample per round, bits set with atomicOr) must give the same words as the default
    one-thread-per-pair kernel."""
    import torch

    from acrobotics_b200 import _native

    g = golden("ccd.npz")
    scene = scene16()
    for k, n in ((_kuka(g), 60_001), (ab.KukaOnRail(), 40_000), (ab.AnthropomorphicArm(), 5_000)):
        qs = torch.as_tensor(random_configs(n, k.ndof, seed=41)).cuda()
        qg = qs + (torch.rand_like(qs) * 2 - 1) * 1.5
        want = k.is_path_in_collision_batch(qs, qg, scene, return_near=True, margin=MARGIN, packed=True)
        want_plain = k.is_path_in_collision_batch(qs, qg, scene, packed=True)
        _native.set_option("swept_persistent", "1")
        try:
            got = k.is_path_in_collision_batch(qs, qg, scene, return_near=True, margin=MARGIN, packed=True)
            plain = k.is_path_in_collision_batch(qs, qg, scene, packed=True)
            # an `out` buffer with stale bits: the words are rewritten, not ORed in
            out = torch.full((n // 32 + 5,), -1, dtype=torch.int32, device="cuda")
            k.is_path_in_collision_batch(qs, qg, scene, out=out)
        finally:
            _native.set_option("swept_persistent", "0")
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        assert torch.equal(plain, want_plain)
        assert torch.equal(out[: (n + 31) // 32], plain) and bool((out[(n + 31) // 32:] == -1).all())


def test_shape_level_seam(gpu):
    """``Shape.is_path_in_collision`` / ``ShapeSoup.is_path_in_collision`` (shapes.py:60-75,
    geometry.py:68-81): the box-level entry point against the scalar FCL restatement, and the
    robot-level verdict recomposed from it the way robot.py:285-317 does."""
    from oracle import fcl_ccd

    rng = np.random.default_rng(8)

    def rand_tf(scale):
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        R *= np.sign(np.linalg.det(R))
        tf = np.eye(4)
        tf[:3, :3], tf[:3, 3] = R, rng.uniform(-scale, scale, 3)
        return tf

    n_hit = 0
    for _ in range(60):
        a, b = ab.Box(*rng.uniform(0.1, 0.6, 3)), ab.Box(*rng.uniform(0.1, 0.6, 3))
        t0, t1, tb = rand_tf(0.6), rand_tf(0.6), rand_tf(0.3)
        want, _, g = fcl_ccd.continuous_collide_boxes(
            [a.dx, a.dy, a.dz], t0[:3, :3], t0[:3, 3], t1[:3, :3], t1[:3, 3],
            [b.dx, b.dy, b.dz], tb[:3, :3], tb[:3, 3])
        if abs(g) > 1e-6:
            assert a.is_path_in_collision(t0, t1, b, tb) == want
            n_hit += want
    assert 5 < n_hit < 55
    # the robot-level query recomposed from the soup-level one (links only: a bare Kuka)
    k = ab.Kuka()
    g = golden("ccd.npz")
    scene = product_scene(g["s3_sides"], g["s3_tfs"])
    for i in range(12):
        qs, qg = g["s3_q_start"][i], g["s3_q_goal"][i]
        fa, fb = k.fk_all_links(qs), k.fk_all_links(qg)
        want = any(l.geometry.is_path_in_collision(fa[j], fb[j], scene) for j, l in enumerate(k.links))
        assert k.is_path_in_collision(qs, qg, scene) == want
