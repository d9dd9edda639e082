"""GPU parity: the CUDA path (through the C ABI, via the host API) against the
oracle and the committed golden vectors.  Tolerances (BASELINE.json):
FK / Jacobians <= 1e-9 abs in FP64 (1e-5 in FP32), IK sets <= 1e-9, collision
booleans bit-exact outside the 1e-6 m near-contact bucket."""
import numpy as np
import pytest
from conftest import golden
from helpers import ROBOTS, kuka_with_tool, product_scene, spec

import acrobotics_b200 as ab
from acrobotics_b200.synthetic import random_configs, readme_path, readme_scene, scene16
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu
FK_TOL = 1e-9
FK_TOL_F32 = 1e-5
MARGIN = 1e-6
IK_TOL = 1e-9


def _assert_ik_parity(sol, ref, valid, what, spec_for_fk=None, T=None):
    """IK solution sets within 1e-9 (BASELINE north star) - no widened tolerance.  Solutions at
    a wrist singularity (|sin q5| < 1e-3) are the one place where the closed form itself is
    ill-conditioned (q4 and q6 are individually determined only to rounding / |sin q5|; numpy
    float64 differs from a long-double evaluation by 1.5e-8 there): they are COUNTED and
    REPORTED, held to the 1e-9 bound on the well-conditioned quantities (q1..q3, q5 and the
    pose they reproduce) and excluded from the per-joint bound on q4 / q6 only."""
    err = np.abs(sol - ref)
    err[~valid] = 0.0
    s5 = np.abs(np.sin(np.where(valid, ref[..., 4], 1.0)))
    singular = valid & (s5 < 1e-3)
    regular = valid & ~singular
    worst_regular = err[regular].max(initial=0.0)
    assert worst_regular <= IK_TOL, f"{what}: worst regular solution {worst_regular:.3e}"
    n_sing = int(singular.sum())
    worst_pose = 0.0
    if n_sing:
        assert err[singular][:, [0, 1, 2]].max() <= IK_TOL, what
        if spec_for_fk is not None:
            idx = np.nonzero(singular)
            pose = no.fk(spec_for_fk, sol[idx])
            worst_pose = float(np.abs(pose - T[idx[0]]).max())
            # the pose error is first order in the q4/q6 rounding times |sin q5| <= 1e-3
            assert worst_pose <= 1e-9, f"{what}: singular solutions reproduce the pose to {worst_pose:.3e}"
    print(f"[ik-parity] {what}: {int(valid.sum())} solutions, worst |dq| {worst_regular:.2e} on the "
          f"regular ones; {n_sing} wrist-singular solutions (|sin q5| < 1e-3) reported separately, "
          f"worst |dq4,dq6| {err[singular].max(initial=0.0):.2e}, worst pose error {worst_pose:.2e}")
    assert singular.mean() < 5e-3, what
    return n_sing


def _assert_collision_parity(gpu_hit, gpu_near, g_oracle, what):
    """Bit-exact outside the margin bucket; the bucket is counted separately."""
    ora_hit = ~(g_oracle > 0)
    ora_near = np.abs(g_oracle) <= MARGIN
    differ = gpu_hit != ora_hit
    outside = differ & ~(ora_near | gpu_near)
    assert not outside.any(), f"{what}: {outside.sum()} verdicts differ outside the margin bucket"
    # the bucket itself must be tiny for random inputs
    assert (ora_near | gpu_near).mean() < 1e-3, what


# ----------------------------------------------------------------------------- FK
@pytest.mark.parametrize("name", ROBOTS)
def test_fk_golden(gpu, name):
    g = golden(f"fk_{name}.npz")
    r = getattr(ab, name)()
    assert np.abs(r.fk_batch(g["q"]) - g["fk"]).max() <= FK_TOL
    assert np.abs(r.fk_batch(g["q"], all_links=True) - g["fk_all_links"]).max() <= FK_TOL
    assert np.abs(r.fk_rpy_batch(g["q"]) - g["fk_rpy"]).max() <= FK_TOL
    assert np.abs(r.fk_batch(g["q"], dtype="f32") - g["fk"]).max() <= FK_TOL_F32
    # scalar API, list input (tests/test_robot.py passes lists)
    T = r.fk(list(g["q"][0]))
    assert T.shape == (4, 4) and np.abs(T - g["fk"][0]).max() <= FK_TOL
    links = r.fk_all_links(g["q"][1])
    assert len(links) == r.ndof and np.abs(links[-1] - g["fk_all_links"][1, -1]).max() <= FK_TOL
    assert np.abs(r.fk_rpy(g["q"][2]) - g["fk_rpy"][2]).max() <= FK_TOL


def test_fk_base_and_tool_tip(gpu):
    g = golden("fk_Kuka_base_tip.npz")
    r = ab.Kuka()
    r.tf_base, r.tf_tool_tip = g["tf_base"], g["tf_tool_tip"]
    assert np.abs(r.fk_batch(g["q"]) - g["fk"]).max() <= FK_TOL
    assert np.abs(r.fk_batch(g["q"], all_links=True) - g["fk_all_links"]).max() <= FK_TOL
    assert np.abs(r.fk_rpy_batch(g["q"]) - g["fk_rpy"]).max() <= FK_TOL


def test_fk_one_million(gpu):
    """BASELINE config 2: 1 M random configs, FP64, pose error vs the oracle."""
    r = ab.Kuka()
    q = random_configs(1_000_000, 6, seed=0)
    T = r.fk_batch(q)
    s = spec(r)
    worst = 0.0
    for lo in range(0, len(q), 100_000):
        worst = max(worst, np.abs(T[lo:lo + 100_000] - no.fk(s, q[lo:lo + 100_000])).max())
    assert worst <= FK_TOL
    # rotation blocks stay orthonormal (size-independent property)
    R = T[:, :3, :3]
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12
    T32 = r.fk_batch(q[:200_000], dtype="f32")
    assert np.abs(T32 - T[:200_000]).max() <= FK_TOL_F32


def test_fk_edge_cases(gpu):
    import torch

    r = ab.Kuka()
    assert r.fk_batch(np.zeros((0, 6))).shape == (0, 4, 4)
    # ragged sizes around the 128-configuration tile / 32-bit mask word
    s = spec(r)
    for n in (1, 31, 33, 127, 129, 257):
        q = random_configs(n, 6, seed=n)
        assert np.abs(r.fk_batch(q) - no.fk(s, q)).max() <= FK_TOL
    # device tensors in -> device tensors out
    qd = torch.from_numpy(random_configs(100, 6, seed=5)).cuda()
    Td = r.fk_batch(qd)
    assert Td.is_cuda and np.abs(Td.cpu().numpy() - no.fk(s, qd.cpu().numpy())).max() <= FK_TOL
    with pytest.raises(ValueError):
        r.fk_batch(np.zeros((4, 5)))
    # link-level transform (link.py:35-58), revolute and prismatic
    rail = ab.KukaOnRail()
    T = rail.links[0].get_link_relative_transform(0.3)
    assert np.abs(T - no.dh_transforms(spec(rail), np.array([[0.3, 0, 0, 0, 0, 0, 0]]))[0, 0]).max() <= FK_TOL


# ------------------------------------------------------------------------ collision
def test_boxbox_seam(gpu):
    g = golden("boxbox.npz")
    for i, expected in enumerate(g["known_hit"]):
        b1, b2 = ab.Box(*g["known_sides1"][i]), ab.Box(*g["known_sides2"][i])
        assert b1.is_in_collision(g["known_tf1"][i], b2, g["known_tf2"][i]) == int(expected)
    from acrobotics_b200.boxes import _pack_boxes, collide_box_pairs

    a = _pack_boxes([ab.Box(*s) for s in g["rand_sides1"]], g["rand_tf1"])
    b = _pack_boxes([ab.Box(*s) for s in g["rand_sides2"]], g["rand_tf2"])
    hit = collide_box_pairs(a, b)
    _, gq = no.sat(0.5 * g["rand_sides1"], g["rand_tf1"][:, :3, :3], g["rand_tf1"][:, :3, 3],
                   0.5 * g["rand_sides2"], g["rand_tf2"][:, :3, :3], g["rand_tf2"][:, :3, 3])
    robust = np.abs(gq) > MARGIN
    assert (hit[robust] == g["rand_hit"][robust]).all()


def test_scene_soup(gpu):
    # reference tests/test_geometry.py:42-59
    from acrobotics_b200.transforms import pose_z

    far = np.eye(4)
    far[0, 3] = 10.0
    s1 = ab.Scene([ab.Box(1, 1, 1)], [np.eye(4)])
    s2 = ab.Scene([ab.Box(1, 1, 2)], [np.eye(4)])
    assert s1.is_in_collision(s2) is True
    assert s1.is_in_collision(s2, tf_self=far) is False
    assert s1.is_in_collision(s2, tf_other=far) is False
    s3 = ab.Scene([ab.Box(1, 1, 1), ab.Box(0.3, 0.2, 0.1)], [pose_z(0.0, -1.0, -1.0, 0.0), -far])
    s4 = ab.Scene([ab.Box(1, 1, 2), ab.Box(0.3, 0.2, 0.1)], [pose_z(np.pi / 4, -2, -2, 0), far])
    assert s3.is_in_collision(s4) is False
    assert s4.is_in_collision(s3) is False


def test_readme_collision_path(gpu):
    robot = ab.Kuka()
    scene = readme_scene()
    out = [robot.is_in_collision(q, scene) for q in readme_path()]
    assert out == [False, False, False, False, True, True, True, True, False, False]
    assert all(isinstance(o, bool) for o in out)
    assert robot.cc_checks == 10


def test_reference_test_robot_cases(gpu):
    g = golden("cc_kuka.npz")
    bot = ab.Kuka()
    gl = [l.geometry for l in bot.links]
    q0 = [0, np.pi / 2, 0, 0, 0, 0]
    assert bot._check_self_collision(bot.fk_all_links(q0), gl) is False
    assert bot._check_self_collision(bot.fk_all_links([0, 1.5, -1.3, 0, -1.5, 0]), gl) is True
    assert bot.is_in_self_collision(q0) is False
    assert bot.is_in_self_collision([0, 1.5, -1.3, 0, -1.5, 0]) is True
    sc1 = product_scene(g["test_robot_sides"], g["test_robot_tfs1"])
    sc2 = product_scene(g["test_robot_sides"], g["test_robot_tfs2"])
    assert bot.is_in_collision(q0, sc1) is True
    assert bot.is_in_collision(q0, sc2) is False


def test_collision_golden_random(gpu):
    g = golden("cc_kuka.npz")
    k = ab.Kuka()
    q = g["rand_q"]
    s = spec(k)
    for scene, key in ((readme_scene(), "rand_hit_readme"), (scene16(), "rand_hit_scene16")):
        hit, near = k.is_in_collision_batch(q, scene, return_near=True)
        plain = k.is_in_collision_batch(q, scene)
        assert (plain == hit).all()  # the margin kernel and the fast kernel agree
        ok = ~near
        assert (hit[ok] == g[key][ok]).all()
        _assert_collision_parity(hit, near, no.collision_margins(s, scene, q), key)
    hit = k.is_in_self_collision_batch(q)
    assert (hit == g["rand_hit_self"]).all()
    k.do_check_self_collision = False
    assert (k.is_in_collision_batch(q, scene16()) == g["rand_hit_scene16_noself"]).all()


def test_collision_tool_base_rail(gpu):
    g = golden("cc_kuka.npz")
    k = kuka_with_tool(g)
    assert np.abs(k.fk_batch(g["tool_q"][:32]) - g["tool_fk"]).max() <= FK_TOL
    assert (k.is_in_collision_batch(g["tool_q"], scene16()) == g["tool_hit_scene16"]).all()
    assert (k.is_in_self_collision_batch(g["tool_q"]) == g["tool_hit_self"]).all()
    rail = ab.KukaOnRail()
    assert (rail.is_in_collision_batch(g["rail_q"], scene16()) == g["rail_hit_scene16"]).all()


def _c_oracle_margins(robot, scene, q, threads=None):
    """hit, g of the plain-C oracle over all rows of q (threads: ctypes releases the GIL)."""
    import concurrent.futures as cf
    import os

    from oracle import c_oracle

    packed = c_oracle.pack(spec(robot), scene)
    c_oracle.load()
    threads = threads or (os.cpu_count() or 1)
    chunks = np.array_split(np.arange(len(q)), threads * 4)
    with cf.ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda idx: c_oracle.fk_cc_packed(packed, q[idx], want_g=True), chunks))
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


def test_collision_one_million_vs_oracle(gpu):
    """>= 1 M random configurations against the oracle, 16-box scene: bit-exact
    outside the 1e-6 m bucket, bucket reported.  All 1 M against the plain-C oracle, the first
    200 k against the vectorised numpy oracle as well."""
    k = ab.Kuka()
    scene = scene16()
    q = random_configs(1_000_000, 6, seed=1)
    hit, near = k.is_in_collision_batch(q, scene, return_near=True)
    assert (k.is_in_collision_batch(q, scene) == hit).all()
    c_hit, c_g = _c_oracle_margins(k, scene, q)
    assert (c_hit == ~(c_g > 0)).all()
    _assert_collision_parity(hit, near, c_g, "scene16 1M vs C oracle")
    print(f"[cc-parity] scene16 1M: 0 mismatches outside the margin, "
          f"{int((np.abs(c_g) <= MARGIN).sum())} configurations inside +-{MARGIN:g} m, "
          f"{int(near.sum())} flagged by the kernel")
    gq = no.collision_margins(spec(k), scene, q[:200_000])
    _assert_collision_parity(hit[:200_000], near[:200_000], gq, "scene16 200k vs numpy oracle")
    assert np.abs(gq - c_g[:200_000]).max() < 1e-12  # the two restatements agree on the margin
    assert 0.5 < hit.mean() < 0.95
    # packed words and booleans describe the same set
    words = k.is_in_collision_batch(q, scene, packed=True)
    assert np.unpackbits(words.view(np.uint8), bitorder="little")[: len(q)].sum() == hit.sum()


def test_collision_bench_inputs_vs_c_oracle(gpu):
    """The benchmark's own generator and seed (bench.py: random_configs_device(n, 6, 1000)):
    the first 1 M configurations of the timed input against the C oracle, self-collision-only
    and README scene included."""
    from acrobotics_b200.synthetic import random_configs_device

    k = ab.Kuka()
    qd = random_configs_device(1 << 20, 6, seed=1000)
    q = qd.cpu().numpy()
    for scene, label in ((scene16(), "scene16"), (readme_scene(), "readme")):
        hit, near = k.is_in_collision_batch(qd, scene, return_near=True)
        hit, near = hit.cpu().numpy(), near.cpu().numpy()
        _, c_g = _c_oracle_margins(k, scene, q)
        _assert_collision_parity(hit, near, c_g, f"bench input {label}")
    hit = k.is_in_self_collision_batch(qd).cpu().numpy()
    _, c_g = _c_oracle_margins(k, None, q)
    assert ((hit == ~(c_g > 0)) | (np.abs(c_g) <= MARGIN)).all()


def test_collision_verbatim_reference_10k(gpu):
    """>= 10 k configurations against the VERBATIM reference's verdicts."""
    g = golden("cc_kuka_10k.npz")
    k = ab.Kuka()
    q = g["q"]
    n = len(q)
    assert n >= 10_000
    for scene, key in ((scene16(), "hit_scene16"), (readme_scene(), "hit_readme")):
        want = np.unpackbits(g[key])[:n].astype(bool)
        hit, near = k.is_in_collision_batch(q, scene, return_near=True)
        assert not near.any(), "a golden configuration sits inside the margin bucket"
        assert (hit == want).all(), key
    want = np.unpackbits(g["hit_self"])[:n].astype(bool)
    assert (k.is_in_self_collision_batch(q) == want).all()
    # scalar API on a few of them
    for i in (0, 17, 4242):
        assert k.is_in_collision(q[i], scene16()) == bool(np.unpackbits(g["hit_scene16"])[i])


def test_collision_properties_and_edges(gpu):
    k = ab.Kuka()
    scene = scene16()
    assert k.is_in_collision_batch(np.zeros((0, 6)), scene).shape == (0,)
    q = random_configs(4097, 6, seed=2)
    full = k.is_in_collision_batch(q, scene)
    # ragged batch sizes: every prefix gives the prefix of the result
    for n in (1, 31, 32, 33, 127, 128, 129, 1000):
        assert (k.is_in_collision_batch(q[:n], scene) == full[:n]).all()
    # joint angles are periodic: q + 2*pi gives the same verdict (q + 2 pi is rounded, i.e. the
    # angle moves by ~4e-16 rad: only a configuration inside the margin bucket may change)
    shifted, near_s = k.is_in_collision_batch(q + 2 * np.pi, scene, return_near=True)
    _, near_0 = k.is_in_collision_batch(q, scene, return_near=True)
    assert ((shifted == full) | near_s | near_0).all()
    print(f"[cc-parity] periodicity: {int((shifted != full).sum())} of {len(q)} verdicts changed, "
          f"all inside the margin bucket ({int((near_s | near_0).sum())} configurations)")
    # permutation equivariance
    perm = np.random.default_rng(0).permutation(len(q))
    assert (k.is_in_collision_batch(q[perm], scene) == full[perm]).all()
    # adding an obstacle can only add collisions; an empty scene leaves self collision
    empty = ab.Scene([], [])
    self_only = k.is_in_self_collision_batch(q)
    assert (k.is_in_collision_batch(q, empty) == self_only).all()
    assert (full | ~self_only).all()
    # FP32 variant: agrees except near contact
    f32 = k.is_in_collision_batch(q, scene, dtype="f32")
    assert (f32 != full).mean() < 2e-3
    with pytest.raises(ValueError):
        k.is_in_collision([0] * 6, None)


# ------------------------------------------------------------------------------- IK
def test_ik_kuka_golden(gpu):
    g = golden("ik_kuka.npz")
    k = ab.Kuka()
    sol, valid = k.ik_batch(g["T"])
    assert (valid.sum(1) == g["count"]).all()
    for i in range(len(sol)):
        assert np.abs(sol[i][valid[i]] - g["sol"][i, : g["count"][i]]).max() <= 1e-9
    assert np.isnan(sol[~valid]).all()
    res = k.ik(g["T"][0])
    assert res.success and len(res.solutions) == 4
    assert np.abs(np.array(res.solutions) - g["sol"][0, :4]).max() <= 1e-9
    k.tf_base, k.tf_tool_tip = g["bt_base"], g["bt_tip"]
    sol, valid = k.ik_batch(g["bt_T"])
    assert (valid.sum(1) == g["bt_count"]).all()
    for i in range(len(sol)):
        assert np.abs(sol[i][valid[i]] - g["bt_sol"][i, : g["bt_count"][i]]).max() <= 1e-9
    k2 = ab.Kuka()
    assert [k2.ik(T).success for T in g["far_T"]] == g["far_success"].tolist()


def test_ik_one_million_vs_oracle(gpu):
    """BASELINE config 4 (IK part): 1 M reachable poses, all branches."""
    k = ab.Kuka()
    s = spec(k)
    q = random_configs(1_000_000, 6, seed=4)
    T = k.fk_batch(q)
    sol, valid = k.ik_batch(T)
    assert valid.any(1).all()
    o_sol, o_valid = no.ik_kuka(s, T[:100_000])
    assert (o_valid == valid[:100_000]).all()
    _assert_ik_parity(sol[:100_000], o_sol, o_valid, "ik 100k", spec_for_fk=s, T=T[:100_000])
    # round trip on the full batch: fk(ik(T)) == T.  The reference asserts 5 decimals on a
    # handful of poses (tests/test_inverse_kinematics.py:120); its elbow snap (|c3| > 1 - 1e-6
    # -> +-1, arm_2.py:41-43) moves q3 by up to 1.4e-3 rad, so over 1 M poses the bound is looser
    for slot in (0, 3, 7):
        m = valid[:, slot]
        err = np.abs(k.fk_batch(sol[m, slot]) - T[m]).reshape(m.sum(), -1).max(axis=1)
        assert err.max() < 1e-3 and np.quantile(err, 0.999) < 1e-9


def test_ik_parts_golden(gpu):
    from acrobotics_b200 import closed_form_ik as cf

    g = golden("ik_parts.npz")
    sol, valid = cf.arm2_batch(g["arm2_p"], 0.18, 0.6, 0.62)
    assert (valid.sum(1) == g["arm2_count"]).all()
    for i in np.nonzero(g["arm2_count"])[0]:
        assert np.abs(sol[i][valid[i]] - g["arm2_sol"][i, : g["arm2_count"][i]]).max() <= 1e-9
    sol, ok = cf.anthropomorphic_batch(g["anth_p"], 1.0, 1.0)
    assert (ok == g["anth_ok"]).all()
    assert np.abs(sol[ok] - g["anth_sol"][ok]).max() <= 1e-9
    sol = cf.wrist_batch(g["wrist_R"], g["wrist_Rb"])
    assert np.abs(sol - g["wrist_sol"]).max() <= 1e-9
    # scalar forms with the reference's signatures
    arm = ab.Arm2(a1=0.18, a2=0.6, a3=0.62)
    i = int(np.nonzero(g["arm2_count"] == 4)[0][0])
    T = np.eye(4)
    T[:3, 3] = g["arm2_p"][i]
    res = arm.ik(T)
    assert res.success and np.abs(np.array(res.solutions) - g["arm2_sol"][i]).max() <= 1e-9
    with pytest.raises(NotImplementedError):
        ab.Robot(arm.links).ik(T)


def test_ik_fused_collision_filter(gpu):
    """BASELINE config 4: IK -> collision filter in one kernel equals IK then K2."""
    k = ab.Kuka()
    scene = scene16()
    q = random_configs(50_000, 6, seed=6)
    T = k.fk_batch(q)
    sol, valid, free = k.ik_filtered_batch(T, scene)
    sol2, valid2 = k.ik_batch(T)
    assert (valid == valid2).all() and np.array_equal(sol, sol2, equal_nan=True)
    flat = sol[valid]
    hit = k.is_in_collision_batch(flat, scene)
    assert (free[valid] == ~hit).all()
    assert not free[~valid].any()
    gq = no.collision_margins(spec(k), scene, flat[:20_000])
    robust = np.abs(gq) > MARGIN
    assert (hit[:20_000][robust] == ~(gq[robust] > 0)).all()
    # unreachable poses (absent branches) mixed in; the small-batch kernel (one fused launch)
    # and the large-batch pipeline (IK kernel -> filtered K2 -> mask combine) agree
    T2 = T[:4001].copy()
    T2[::5, :3, 3] *= 3.0
    sol_a, valid_a, free_a = k.ik_filtered_batch(T2, scene)
    assert (valid_a[::5] == False).any() and valid_a.any()  # noqa: E712
    assert not free_a[~valid_a].any()
    for lo in range(0, 4001, 500):
        sol_b, valid_b, free_b = k.ik_filtered_batch(T2[lo:lo + 500], scene)
        assert (valid_b == valid_a[lo:lo + 500]).all() and (free_b == free_a[lo:lo + 500]).all()


def test_kuka_on_rail_ik(gpu):
    rail = ab.KukaOnRail()
    q = np.array([0.4, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6])
    res = rail.ik(rail.fk(q), q[0])
    assert res.success
    assert min(np.abs(np.array(s) - q).max() for s in res.solutions) < 1e-6
    # batched, one rail position: every valid branch reproduces its pose
    qs = random_configs(2000, 7, seed=19)
    qs[:, 0] = 0.4
    T = rail.fk_batch(qs)
    sol, valid = rail.ik_batch(T, 0.4)
    assert sol.shape == (2000, 8, 7) and valid.any(axis=1).all()
    back = rail.fk_batch(sol[valid])
    assert np.abs(back - np.repeat(T, 8, axis=0).reshape(2000, 8, 4, 4)[valid]).max() < 1e-9


# ------------------------------------------------------------------------ Jacobians
@pytest.mark.parametrize("name", ["Kuka", "PlanarArm", "KukaOnRail"])
def test_jacobians(gpu, name):
    g = golden("jac.npz")
    r = getattr(ab, name)()
    q = g[f"{name}_q"]
    Jp = r.jacobian_position_batch(q)
    Jr = r.jacobian_rpy_batch(q)
    s = spec(r)
    assert np.abs(Jp - no.jacobian_position(s, q)).max() <= 1e-9
    assert np.abs(Jr - no.jacobian_rpy(s, q)).max() <= 1e-9
    # the reference test's own criterion (tests/test_jacobians.py:74-95)
    assert np.abs(Jp - g[f"{name}_jpos"]).max() < 1e-7
    assert np.abs(Jr[:, 3:] - g[f"{name}_jrpy"][:, 3:]).max() < 1e-5
    assert r.jacobian_position(q[0]).shape == (3, r.ndof)
    assert r.jacobian_rpy(q[0]).shape == (6, r.ndof)
    big = random_configs(100_000, r.ndof, seed=9) * 0.3
    assert np.abs(r.jacobian_rpy_batch(big) - no.jacobian_rpy(s, big)).max() <= 1e-9


# ----------------------------------------------------------------------- path sweep
def test_path_sweep_golden(gpu):
    g = golden("path.npz")
    k = ab.Kuka()
    hit = k.is_path_in_collision_discrete_batch(g["q_start"], g["q_goal"], readme_scene(), 0.1)
    assert (hit == g["hit_readme"]).all()
    hit = k.is_path_in_collision_discrete_batch(g["q_start"], g["q_goal"], scene16(), 0.05)
    assert (hit == g["hit_scene16"]).all()
    assert k.is_path_in_collision_discrete(g["q_start"][5], g["q_goal"][5], readme_scene()) == bool(
        g["hit_readme"][5]
    )
    # a sweep equals the OR of K2 over its interpolants
    for e in (7, 11, 23):
        P = np.array(k._linear_interpolation_path(g["q_start"][e], g["q_goal"][e], 0.05))
        assert bool(k.is_in_collision_batch(P, scene16()).any()) == bool(g["hit_scene16"][e])


def test_path_sweep_large(gpu):
    k = ab.Kuka()
    scene = scene16()
    rng = np.random.default_rng(11)
    qs = rng.uniform(-np.pi, np.pi, (20_000, 6))
    qg = qs + rng.uniform(-0.4, 0.4, qs.shape)
    hit = k.is_path_in_collision_discrete_batch(qs, qg, scene, 0.1)
    # oracle with margin bookkeeping: a segment is robust when none of its interpolants is
    # within 1e-6 m of contact; robust segments must agree exactly, the rest is counted
    m = 2000
    s_ = spec(k)
    robust = np.ones(m, bool)
    ora = np.zeros(m, bool)
    for e in range(m):
        P = np.array(no.interpolation_path(qs[e], qg[e], 0.1)).reshape(-1, 6)
        if len(P):
            gq = no.collision_margins(s_, scene, P)
            ora[e] = bool((gq <= 0).any())
            robust[e] = bool((np.abs(gq) > MARGIN).all())
    assert (hit[:m][robust] == ora[robust]).all()
    print(f"[cc-parity] path sweep: {int(robust.sum())} robust segments agree exactly, "
          f"{int((~robust).sum())} segments have an interpolant inside the margin bucket "
          f"({int((hit[:m][~robust] != ora[~robust]).sum())} of them differ)")
    assert (~robust).mean() < 5e-3
    # an endpoint in collision implies the sweep is in collision
    start_hit = k.is_in_collision_batch(qs, scene)
    assert (hit | ~start_hit).all()


# ------------------------------------------------------------------- the boundary from C
def test_c_program_on_the_abi(gpu, tmp_path):
    """examples/c_abi_smoke.c: a plain C program (no Python objects, no torch) builds the Kuka and
    the README scene through include/acro_b200.h and reproduces README.md:94-96."""
    import os
    import shutil
    import subprocess

    from conftest import ROOT

    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(f"{cuda}/include/cuda_runtime_api.h"):
        pytest.skip("gcc / CUDA toolkit not available")
    exe = tmp_path / "c_abi_smoke"
    lib_dir = os.path.join(ROOT, "acrobotics_b200")
    subprocess.run(
        [gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", f"{cuda}/include",
         os.path.join(ROOT, "examples", "c_abi_smoke.c"), "-L", lib_dir, "-lacro_b200",
         "-L", f"{cuda}/lib64", "-lcudart", "-lm", "-o", str(exe)], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=f"{lib_dir}:{cuda}/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([str(exe)], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "F F F F T T T T F F" in out.stdout
