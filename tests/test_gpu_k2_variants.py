"""K2 kernel variants: the filtered production kernel (FP32 filter + candidate grid +
exact FP64 fix-up) must return exactly the verdicts of the plain FP64 kernel - on random
batches, on inputs built to sit on the filter's ambiguity band (aligned boxes, grazing
contact), outside the candidate grid, and for joint values the FP32 chain cannot follow."""
import numpy as np
import pytest
from conftest import golden
from helpers import kuka_with_tool, spec

import acrobotics_b200 as ab
from acrobotics_b200 import _native
from acrobotics_b200.synthetic import random_configs, readme_scene, scene16
from acrobotics_b200.transforms import translation
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu


def _with_variant(name, fn):
    _native.set_option("k2_variant", name)
    try:
        return fn()
    finally:
        _native.set_option("k2_variant", "filter")


def _all_variants_agree(robot, q, scene, self_only=False):
    if scene is not None and not self_only:
        scene.prepare(robot)  # batches below 64 K only take the filtered kernel with a plan
    call = (lambda: robot.is_in_self_collision_batch(q)) if self_only else (
        lambda: robot.is_in_collision_batch(q, scene))
    ref = _with_variant("simple", call)
    got = _with_variant("filter", call)
    assert (got == ref).all(), f"{int((got != ref).sum())} of {len(q)} verdicts differ"
    return ref


def test_filter_equals_fp64_on_random_batches(gpu):
    g = golden("cc_kuka.npz")
    q = random_configs(2_000_000, 6, seed=11)
    k = ab.Kuka()
    for scene in (scene16(), readme_scene(), scene16(seed=5, n_boxes=40), ab.Scene([], [])):
        hit = _all_variants_agree(k, q, scene)
        assert 0.05 < hit.mean() < 0.99
    _all_variants_agree(k, q, None, self_only=True)
    wave = _with_variant("wave", lambda: k.is_in_collision_batch(q, scene16()))
    assert (wave == k.is_in_collision_batch(q, scene16())).all()
    # tool + base geometry + moved base, prismatic first joint
    _all_variants_agree(kuka_with_tool(g), q[:500_000], scene16())
    _all_variants_agree(ab.KukaOnRail(), random_configs(500_000, 7, seed=16), scene16())
    k.do_check_self_collision = False
    _all_variants_agree(k, q[:500_000], scene16())


def test_filter_vs_oracle(gpu):
    k = ab.Kuka()
    scene = scene16()
    q = random_configs(300_000, 6, seed=12)
    hit = k.is_in_collision_batch(q, scene)
    gq = no.collision_margins(spec(k), scene, q)
    robust = np.abs(gq) > 1e-6
    assert (hit[robust] == ~(gq[robust] > 0)).all()
    assert (~robust).mean() < 1e-3


def test_filter_ambiguity_band(gpu):
    """Configurations engineered to land inside the filter tolerance: axis-aligned boxes
    (degenerate edge axes), links grazing an obstacle face within 1e-7 .. 1e-4 m."""
    k = ab.Kuka()
    rng = np.random.default_rng(3)
    n = 8192
    # joint angles on multiples of pi/2 keep every link box axis-aligned with the scene
    q_aligned = rng.integers(-2, 3, (n, 6)) * (np.pi / 2)
    q_aligned[::2] += rng.normal(0, 1e-7, (n // 2, 6))
    _all_variants_agree(k, q_aligned, readme_scene())
    # a wall whose face sits within +-gap of link 2's box face for q = 0 pose:
    # sweep the wall position through contact
    T = k.fk_batch(np.zeros((1, 6)), all_links=True)[0]
    c = T[1] @ np.array([-0.3, 0, -0.05, 1.0])  # centre of link-1 box
    for gap in (1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 0.0, -1e-7, -1e-5):
        wall = ab.Scene([ab.Box(0.2, 2.0, 2.0)],
                        [translation(c[0] + 0.4 + 0.1 + gap, c[1], c[2])])
        qs = np.zeros((n, 6))
        qs[:, 3:] = rng.uniform(-3, 3, (n, 3))  # wrist moves, link 1 does not
        qs[:, 0] = rng.normal(0, 1e-6, n)
        _all_variants_agree(k, qs, wall)


def test_filter_out_of_range_inputs(gpu):
    k = ab.Kuka()
    scene = scene16()
    n = 8192
    q = random_configs(n, 6, seed=13)
    q[::7] *= 1e3          # large angles: FP64 argument reduction still exact
    q[5::11] *= 1e9        # beyond the filter's range: decided by the FP64 pass
    q[3::13, 2] = np.nan   # NaN joints: whatever the FP64 kernel says
    q[1::17, 4] = np.inf
    _all_variants_agree(k, q, scene)
    # rail far outside the candidate grid (prismatic travel beyond the covered range)
    rail = ab.KukaOnRail()
    qr = random_configs(n, 7, seed=14)
    qr[:, 0] *= 3.0
    _all_variants_agree(rail, qr, scene)
    # a scene that is far away / huge compared with the robot
    far = ab.Scene([ab.Box(50, 50, 1.0), ab.Box(1, 1, 1)],
                   [translation(0, 0, -0.55), translation(30, 0, 0)])
    _all_variants_agree(k, q, far)


def test_grid_resolution_does_not_change_verdicts(gpu):
    k = ab.Kuka()
    q = random_configs(200_000, 6, seed=15)
    ref = _with_variant("simple", lambda: k.is_in_collision_batch(q, scene16()))
    try:
        for dim in (8, 32, 100):
            _native.set_option("k2_grid", dim)
            assert (k.is_in_collision_batch(q, scene16()) == ref).all()
    finally:
        _native.set_option("k2_grid", 96)


def test_filtered_path_sweep_equals_fp64(gpu):
    """K6: lane-refill filter kernel + FP64 pass over ambiguous segments == warp-per-segment
    FP64 kernel, for short (planner edge), long, empty and mixed segments."""
    k = ab.Kuka()
    rng = np.random.default_rng(21)
    n = 60_000
    qs = rng.uniform(-np.pi, np.pi, (n, 6))
    cases = {
        "short": qs + rng.uniform(-0.15, 0.15, qs.shape),
        "long": rng.uniform(-np.pi, np.pi, (n, 6)),
        "mixed": np.where(rng.random((n, 1)) < 0.5, qs + rng.uniform(-0.05, 0.05, qs.shape),
                          rng.uniform(-np.pi, np.pi, (n, 6))),
    }
    cases["mixed"][::97] = qs[::97]  # zero-length segments: empty path, never in collision
    for name, qg in cases.items():
        for scene, step in ((scene16(), 0.1), (readme_scene(), 0.25)):
            call = lambda: k.is_path_in_collision_discrete_batch(qs, qg, scene, step)  # noqa: E731
            ref = _with_variant("simple", call)
            got = _with_variant("filter", call)
            assert (got == ref).all(), f"{name}: {int((got != ref).sum())} segments differ"
            if name == "mixed":
                assert not got[::97].any()
    # free space: a scene nobody touches -> nothing collides except through self collision
    far = ab.Scene([ab.Box(1, 1, 1)], [translation(30, 0, 0)])
    qg = qs + rng.uniform(-0.3, 0.3, qs.shape)
    call = lambda: k.is_path_in_collision_discrete_batch(qs, qg, far, 0.1)  # noqa: E731
    assert (_with_variant("simple", call) == _with_variant("filter", call)).all()


def test_filter_tolerance_covers_measured_fp32_error(gpu):
    """The filter is exact as long as |s_fp32 - s_fp64| <= delta for every SAT quantity.  Measure
    the left-hand side over 4 M configurations x all link-box x scene-box pairs x 15 axes and
    demand a factor 10 of head-room (typically it is ~60)."""
    import ctypes

    import torch

    lib = _native.load()
    for robot, ndof, scene, scale in ((ab.Kuka(), 6, scene16(), 1.0),
                                      (ab.Kuka(), 6, scene16(seed=3, n_boxes=40), 1.0),
                                      (ab.KukaOnRail(), 7, scene16(), 1.0),
                                      (ab.Kuka(), 6, scene16(), 40.0)):
        q = torch.from_numpy(random_configs(4_000_000 if ndof == 6 else 1_000_000, ndof, seed=31)
                             * scale).cuda()
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        delta = ctypes.c_double()
        _native.check(lib.acro_filter_tolerance(robot._handle(), scene.native_handle(),
                                                ctypes.byref(delta)))
        _native.check(lib.acro_measure_filter_error(robot._handle(), scene.native_handle(),
                                                    q.data_ptr(), q.shape[0], out.data_ptr(), None))
        worst_s, worst_p = out.cpu().tolist()
        print(f"delta {delta.value:.3e}  max|s32-s64| {worst_s:.3e}  max centre error {worst_p:.3e}")
        assert 0 < worst_s < delta.value / 10
        assert worst_p < delta.value / 10


def _measure(robot, scene, q):
    import ctypes

    import torch

    lib = _native.load()
    qd = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    out = torch.zeros(2, dtype=torch.float64, device="cuda")
    delta = ctypes.c_double()
    _native.check(lib.acro_filter_tolerance(robot._handle(), scene.native_handle(), ctypes.byref(delta)))
    _native.check(lib.acro_measure_filter_error(robot._handle(), scene.native_handle(), qd.data_ptr(),
                                                qd.shape[0], out.data_ptr(), None))
    worst_s, worst_p = out.cpu().tolist()
    return delta.value, worst_s, worst_p


def test_filter_tolerance_is_a_bound_at_the_worst_corners(gpu):
    """The tolerance is a derived running error bound (abi.cu: derive_filter_tolerance), so it
    must hold - with room - where FP32 is worst: an 8-joint chain with 2 m links, a base 50 m
    from the origin, 1 mm plates, scene boxes whose edges are parallel to the robot's to within
    1e-7 rad, prismatic travel at its limits, joint angles of 1e6 rad.  For each case: measured
    max |s_fp32 - s_fp64| over all link-box x scene-box pairs x 15 axes < delta, and the
    filtered verdicts equal the all-FP64 kernel's."""
    rng = np.random.default_rng(77)
    cases = []
    # (1) eight revolute joints, a and d up to 2 m
    links = []
    for i in range(8):
        dh = ab.DHLink(float(rng.uniform(0.5, 2.0)), float(rng.choice([0, np.pi / 2, -np.pi / 2, 0.7])),
                       float(rng.uniform(0.0, 2.0)), 0.0)
        links.append(ab.Link(dh, ab.JointType.revolute,
                             ab.Scene([ab.Box(*rng.uniform(0.2, 1.5, 3))], [translation(*rng.uniform(-0.5, 0.5, 3))])))
    long_arm = ab.Robot(links)
    big_scene = scene16(seed=9, n_boxes=32)
    for t in big_scene.tf_s:
        t[:3, 3] *= 6.0
    cases.append(("8 dof, 2 m links", long_arm, big_scene, rng.uniform(-np.pi, np.pi, (400_000, 8))))
    # (2) base 50 m from the origin, scene moved along
    far = ab.Kuka()
    shift = np.array([30.0, -38.0, 12.0])  # |shift| = 49.9
    far.tf_base = translation(*shift)
    far_scene = scene16()
    for t in far_scene.tf_s:
        t[:3, 3] += shift
    cases.append(("base at 50 m", far, far_scene, random_configs(400_000, 6, seed=1)))
    # (3) 1 mm plates: in the scene and on the robot
    thin = ab.Kuka()
    thin.links[2].geometry.s[0] = ab.Box(0.4, 0.3, 0.001)
    plates = ab.Scene([ab.Box(2, 2, 0.001), ab.Box(0.001, 1.0, 1.0), ab.Box(0.6, 0.001, 0.8)],
                      [translation(0, 0, -0.2), translation(0.7, 0, 0.5), translation(0, -0.6, 0.6)])
    cases.append(("1 mm plates", thin, plates, random_configs(400_000, 6, seed=2)))
    # (4) near-parallel edges: scene boxes in link frames of sampled configurations, turned by
    #     1e-7 rad, and configurations around those samples
    k = ab.Kuka()
    q0 = random_configs(12, 6, seed=3)
    frames = k.fk_batch(q0, all_links=True)
    tfs = []
    for i in range(12):
        tf = frames[i, i % 6].copy()
        a = 1e-7
        tf[:3, :3] = tf[:3, :3] @ np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        tf[:3, 3] += tf[:3, :3] @ np.array([0.05, 0.25, 0.0])
        tfs.append(tf)
    par_scene = ab.Scene([ab.Box(0.3, 0.3, 0.3)] * 12, tfs)
    qn = np.repeat(q0, 30_000, axis=0) + rng.normal(0, 1e-3, (360_000, 6))
    cases.append(("near-parallel edges", k, par_scene, qn))
    # (5) prismatic travel at +-q_lin_max (4 m) and beyond (those go to the FP64 pass)
    rail = ab.KukaOnRail()
    qr = random_configs(300_000, 7, seed=4)
    qr[:, 0] = rng.choice([-4.0, 4.0, -3.999, 3.999, 4.5, -7.0], len(qr))
    cases.append(("rail at its limits", rail, scene16(), qr))
    # (6) huge joint angles (revolute |q| up to 1e6)
    cases.append(("|q| ~ 1e6", ab.Kuka(), scene16(), random_configs(300_000, 6, seed=5) * 3e5))
    for name, robot, scene, q in cases:
        qm = q if name != "rail at its limits" else q[np.abs(q[:, 0]) <= 4.0]
        delta, worst_s, worst_p = _measure(robot, scene, qm)
        print(f"[filter-bound] {name:22s} delta {delta:.3e}  max|s32-s64| {worst_s:.3e}  "
              f"(head-room {delta / max(worst_s, 1e-30):5.1f}x)  centre error {worst_p:.3e}")
        assert 0 < worst_s < delta and worst_p < delta, name
        _all_variants_agree(robot, q, scene)


def test_margin_mode_through_the_filter(gpu):
    """return_near=True: verdicts and the +-1e-6 m bucket from the filtered kernel (filter run
    with tolerance delta + margin, FP64 margin evaluation of the flagged configurations) equal
    the all-FP64 margin kernel."""
    k = ab.Kuka()
    rng = np.random.default_rng(5)
    q = random_configs(1_000_000, 6, seed=41)
    for scene in (scene16(), readme_scene()):
        call = lambda: k.is_in_collision_batch(q, scene, return_near=True)  # noqa: E731
        hit_s, near_s = _with_variant("simple", call)
        hit_f, near_f = _with_variant("filter", call)
        assert (hit_s == hit_f).all() and (near_s == near_f).all()
    # grazing contact: a wall face within +-gap of link 1's box face
    T = k.fk_batch(np.zeros((1, 6)), all_links=True)[0]
    c = T[1] @ np.array([-0.3, 0, -0.05, 1.0])
    n = 8192
    total_near = 0
    for gap in (1e-5, 2e-6, 5e-7, 0.0, -5e-7, -2e-6):
        wall = ab.Scene([ab.Box(0.2, 2.0, 2.0)], [translation(c[0] + 0.4 + 0.1 + gap, c[1], c[2])])
        wall.prepare(k)
        qs = np.zeros((n, 6))
        qs[:, 3:] = rng.uniform(-3, 3, (n, 3))
        call = lambda: k.is_in_collision_batch(qs, wall, return_near=True)  # noqa: E731
        hit_s, near_s = _with_variant("simple", call)
        hit_f, near_f = _with_variant("filter", call)
        assert (hit_s == hit_f).all() and (near_s == near_f).all(), gap
        total_near += int(near_f.sum())
    assert total_near > 0
    # a margin wider than the filter tolerance falls back to the FP64 kernel (same answers)
    a = k.is_in_collision_batch(q[:100_000], scene16(), return_near=True, margin=1e-3)
    b = _with_variant("simple", lambda: k.is_in_collision_batch(q[:100_000], scene16(),
                                                                 return_near=True, margin=1e-3))
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[1].sum() > 0


def test_full_size_properties(gpu):
    """BASELINE config 3 at full size (100 M configurations, device resident): the filtered
    kernel and the all-FP64 wavefront kernel return the same 12.5 MB bit mask, shards
    concatenate to the whole, joint angles are 2*pi periodic, and the host-buffer pipeline
    agrees with the device path."""
    import torch

    from acrobotics_b200.synthetic import random_configs_device

    k = ab.Kuka()
    scene = scene16()
    n = 100_000_000
    q = random_configs_device(n, 6, seed=1000)
    words = k.is_in_collision_batch(q, scene, packed=True)
    assert words.shape == ((n + 31) // 32,)
    wave = _with_variant("wave", lambda: k.is_in_collision_batch(q, scene, packed=True))
    assert torch.equal(words, wave)
    del wave
    rate = float(k.is_in_collision_batch(q[: 1 << 22], scene).float().mean())
    assert 0.68 < rate < 0.74
    # shards (cut on a multiple of 32) concatenate to the whole
    cut = 32 * 1_234_567
    a = k.is_in_collision_batch(q[:cut], scene, packed=True)
    b = k.is_in_collision_batch(q[cut:], scene, packed=True)
    assert torch.equal(torch.cat([a, b]), words)
    del a, b
    # periodicity on a 4 M slice: q + 2 pi rounds differently, verdicts differ only at contact
    sl = q[: 1 << 22]
    h0 = k.is_in_collision_batch(sl, scene)
    h1 = k.is_in_collision_batch(sl + 2 * np.pi, scene)
    assert float((h0 != h1).float().mean()) < 1e-5
    # host-buffer pipeline (chunked H2D / kernel / D2H) == device path on an 8 M prefix
    m = 8_000_000
    lib = _native.load()
    q_host = q[:m].cpu().pin_memory()
    mask_host = torch.zeros((m + 31) // 32, dtype=torch.int32).pin_memory()
    _native.check(lib.acro_fk_cc_f64_host(k._handle(), scene.native_handle(), q_host.data_ptr(), m,
                                          mask_host.data_ptr(), 1 << 20))
    assert torch.equal(mask_host.cuda(), words[: (m + 31) // 32])


def test_split_pass_does_not_change_verdicts(gpu):
    """The two-pass schedule (regrouping undecided configurations after slot k) is only a
    schedule: every split point gives the verdicts of the FP64 kernel."""
    g = golden("cc_kuka.npz")
    q = random_configs(700_001, 6, seed=51)
    cases = [(ab.Kuka(), q, scene16()), (kuka_with_tool(g), q[:300_000], scene16()),
             (ab.KukaOnRail(), random_configs(300_000, 7, seed=52), scene16()),
             (ab.Kuka(), q[:300_000], readme_scene())]
    try:
        for robot, qq, scene in cases:
            ref = _with_variant("simple", lambda: robot.is_in_collision_batch(qq, scene))
            for split in ("auto", 0, 1, 2, 3, 5):
                _native.set_option("k2_split", split)
                got = robot.is_in_collision_batch(qq, scene)
                assert (got == ref).all(), (type(robot).__name__, split, int((got != ref).sum()))
                hit, near = robot.is_in_collision_batch(qq[:100_000], scene, return_near=True)
                assert (hit == ref[:100_000]).all()
    finally:
        _native.set_option("k2_split", "auto")


def test_fp32_input_takes_the_exact_path(gpu):
    """Large FP32-input batches: verdicts of the FP64 kernel at the float-valued angles."""
    k = ab.Kuka()
    scene = scene16()
    q = random_configs(300_000, 6, seed=61)
    q32 = q.astype(np.float32)
    got = k.is_in_collision_batch(q32, scene, dtype="f32")
    ref = _with_variant("simple", lambda: k.is_in_collision_batch(q32.astype(np.float64), scene))
    assert (got == ref).all()
    assert (got != k.is_in_collision_batch(q, scene)).mean() < 1e-4  # only the rounding of q


def test_scene_size_limits(gpu):
    """The largest supported scene (64 boxes: one uint64 candidate word per grid cell) and the
    error beyond it."""
    k = ab.Kuka()
    q = random_configs(100_000, 6, seed=71)
    big = scene16(seed=7, n_boxes=64)
    _all_variants_agree(k, q, big)
    with pytest.raises(ValueError):
        k.is_in_collision_batch(q[:10], scene16(seed=7, n_boxes=65))


def test_moving_base_reuses_and_evicts_plans(gpu):
    """A robot whose base moves through one scene: every base pose gets its own candidate grid,
    a scene keeps at most eight of them (oldest unused one evicted); verdicts stay exact."""
    k = ab.Kuka()
    scene = scene16()
    q = random_configs(20_000, 6, seed=81)
    for step in range(12):
        k.tf_base = translation(0.05 * step, -0.03 * step, 0.01 * step)
        _all_variants_agree(k, q, scene)
    k.tf_base = translation(0, 0, 0)
    _all_variants_agree(k, q, scene)  # the first plan was evicted and is rebuilt


def test_small_batches_without_a_plan_use_the_fp64_kernel(gpu):
    """No candidate grid is built for a batch below 64 K configurations (it would cost more
    than the batch); with a prepared plan the same batch takes the filtered kernel.  Same
    verdicts either way, and the launch counter tells which kernel ran."""
    k = ab.Kuka()
    q = random_configs(10_000, 6, seed=91)
    scene = scene16(seed=11)
    before = _native.launch_count()
    a = k.is_in_collision_batch(q, scene)
    assert _native.launch_count() - before == 1  # fk_cc_kernel<double>
    scene.prepare(k)
    before = _native.launch_count()
    b = k.is_in_collision_batch(q, scene)
    assert _native.launch_count() - before == 2  # filter + exact pass
    assert (a == b).all()


def _random_rotation(rng):
    quat = rng.normal(size=4)
    w, x, y, z = quat / np.linalg.norm(quat)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _random_robot(rng):
    """A random serial chain: 3-8 joints (revolute / prismatic), 0-3 boxes per link with
    arbitrary offsets, maybe tool and base geometry, a moved base, a random collision matrix."""
    ndof = int(rng.integers(3, 9))
    links = []
    for i in range(ndof):
        dh = ab.DHLink(float(rng.uniform(0, 0.5)), float(rng.choice([0, np.pi / 2, -np.pi / 2, 0.3])),
                       float(rng.uniform(0, 0.4)), float(rng.uniform(-1, 1)))
        jt = ab.JointType.prismatic if rng.random() < 0.2 else ab.JointType.revolute
        shapes, tfs = [], []
        for _ in range(int(rng.integers(0, 4)) if i else 1):
            tf = np.eye(4)
            if rng.random() < 0.5:
                tf[:3, :3] = _random_rotation(rng)
            tf[:3, 3] = rng.uniform(-0.2, 0.2, 3)
            shapes.append(ab.Box(*rng.uniform(0.05, 0.4, 3)))
            tfs.append(tf)
        links.append(ab.Link(dh, jt, ab.Scene(shapes, tfs)))
    robot = ab.Robot(links)
    if rng.random() < 0.5:
        tf = np.eye(4)
        tf[:3, :3] = _random_rotation(rng)
        tf[:3, 3] = rng.uniform(-0.1, 0.1, 3)
        tip = np.eye(4)
        tip[:3, 3] = [0, 0, 0.1]
        robot.tool = ab.Tool([ab.Box(0.1, 0.05, 0.2), ab.Box(0.05, 0.05, 0.05)],
                             [tf, np.eye(4)], tip)
    if rng.random() < 0.5:
        robot.geometry_base = ab.Scene([ab.Box(0.4, 0.4, 0.2)], [translation(0, 0, -0.1)])
    if rng.random() < 0.5:
        base = np.eye(4)
        base[:3, :3] = _random_rotation(rng)
        base[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        robot.tf_base = base
    cm = rng.random((ndof, ndof)) < 0.4
    robot.collision_matrix = np.triu(cm, 2) | np.triu(cm, 2).T
    robot.do_check_self_collision = bool(rng.random() < 0.8)
    return robot


def test_random_robots_and_scenes(gpu):
    """Fuzz: the filtered kernel (candidate grid, two passes, FP32 filter, FP64 pass) against
    the plain FP64 kernel on random chains, geometries and scenes - exact agreement - and the
    plain kernel against the oracle on a prefix."""
    rng = np.random.default_rng(2024)
    for trial in range(12):
        robot = _random_robot(rng)
        scene = scene16(seed=100 + trial, n_boxes=int(rng.integers(1, 50)))
        q = rng.uniform(-np.pi, np.pi, (150_000, robot.ndof))
        ref = _all_variants_agree(robot, q, scene)
        got_self = _with_variant("filter", lambda: robot.is_in_self_collision_batch(q))
        ref_self = _with_variant("simple", lambda: robot.is_in_self_collision_batch(q))
        assert (got_self == ref_self).all(), trial
        g = no.collision_margins(spec(robot), scene, q[:3000])
        robust = np.abs(g) > 1e-6
        assert (ref[:3000][robust] == ~(g[robust] > 0)).all(), trial
        # the path sweep on the same robot
        qg = q[:20_000] + rng.uniform(-0.3, 0.3, (20_000, robot.ndof))
        call = lambda: robot.is_path_in_collision_discrete_batch(q[:20_000], qg, scene, 0.1)  # noqa: E731
        assert (_with_variant("simple", call) == _with_variant("filter", call)).all(), trial


def test_concurrent_callers(gpu):
    """Handles are immutable and the scratch of a call is stream ordered: several host threads
    (ctypes drops the GIL) check different batches against the same robot / scene handles on
    their own streams and get what a single caller gets."""
    import threading

    import torch

    k = ab.Kuka()
    scene = scene16()
    scene.prepare(k)
    k._handle()  # build the shared robot handle before the threads start
    batches = [torch.from_numpy(random_configs(200_000 + 1000 * i, 6, seed=100 + i)).cuda()
               for i in range(6)]
    expected = [k.is_in_collision_batch(b, scene, packed=True).clone() for b in batches]
    torch.cuda.synchronize()
    results = [None] * len(batches)
    errors = []

    def worker(i):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(5):
                    out = k.is_in_collision_batch(batches[i], scene, packed=True)
                stream.synchronize()
            results[i] = out
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(batches))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for got, exp in zip(results, expected):
        assert torch.equal(got, exp)


def test_many_saved_poses_fall_back_to_small_ctas(gpu):
    """The filter kernel is instantiated for 4, 9 and 14 warps per CTA and the launcher takes
    the size with the most resident warps.  A robot whose self pairs keep many box poses
    (8 links x 2 boxes, every pair of links two or more apart) needs ~25 KB of shared memory
    per warp: only the small CTAs fit, and the verdicts still equal the plain FP64 kernel's."""
    rng = np.random.default_rng(77)
    links = []
    for i in range(8):
        dh = ab.DHLink(0.25, float([np.pi / 2, 0.0, -np.pi / 2, 0.3][i % 4]), 0.1, 0.0)
        shapes = [ab.Box(0.2, 0.08, 0.08), ab.Box(0.08, 0.08, 0.15)]
        tfs = [translation(-0.1, 0, 0), translation(0, 0, 0.05)]
        links.append(ab.Link(dh, ab.JointType.revolute, ab.Scene(shapes, tfs)))
    robot = ab.Robot(links)
    cm = np.ones((8, 8), dtype=bool)
    robot.collision_matrix = np.triu(cm, 2) | np.triu(cm, 2).T
    robot.do_check_self_collision = True
    q = rng.uniform(-np.pi, np.pi, (200_000, 8))
    for scene in (scene16(), scene16(seed=9, n_boxes=40)):
        hit = _all_variants_agree(robot, q, scene)
        assert 0.05 < hit.mean() < 1.0
    _all_variants_agree(robot, q, None, self_only=True)
