"""Generate ``tests/golden/*.npz`` by running the VERBATIM reference
(``/root/reference/src/acrobotics`` through ``oracle.ref_shim``) on seeded inputs
(TEST INFRASTRUCTURE; run in the build container only:
``python -m oracle.make_golden``).

What the fixtures pin
---------------------
fk_<robot>.npz   q, fk, fk_all_links, fk_rpy for every example robot, plus a Kuka
                 with a base transform and a tool tip (reference robot.py:59-91).
cc_kuka.npz      Robot.is_in_collision / is_in_self_collision verdicts
                 (robot.py:206-273): README path, tests/test_robot.py cases,
                 random configurations against the README scene and the
                 16-box benchmark scene, a Kuka with the box-only ``torch2`` tool
                 and a base geometry, KukaOnRail.
ik_kuka.npz      Kuka.ik solution lists (robot_examples.py:188-218), README vector.
ik_parts.npz     arm_2 / spherical_wrist / anthropomorphic_arm ik.
path.npz         _linear_interpolation_path + is_path_in_collision_discrete.
boxbox.npz       Shape.is_in_collision known answers (tests/test_shapes.py:128-142,
                 tests/test_geometry.py:42-59) and random pairs.
jac.npz          central-difference Jacobians of the verbatim fk / fk_rpy (the
                 method of reference tests/test_jacobians.py; casadi is absent).
cc_kuka_10k.npz  SURVEY 8(d)'s ">= 10 k against the verbatim path": 10 000 random Kuka
                 configurations against the 16-box benchmark scene and the README scene
                 plus self-collision verdicts, all through the reference's own
                 Robot.is_in_collision (``python -m oracle.make_golden cc10k``).
polyhedron.npz   Box.get_polyhedron (A, b) of random boxes (shapes.py:163-172): pins
                 oracle/lp_boxbox.py, the FCL-independent Box-Box ground truth
                 (``python -m oracle.make_golden polyhedron``).

The Box-Box predicate inside those verdicts is oracle/fcl_boxbox.py (python-fcl is
not installable here); everything else is the reference's own code.
"""
import os
import sys

import numpy as np

from oracle import ref_shim
from oracle.ref_shim import pose_x, pose_z, translation

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))
from acrobotics_b200.synthetic import scene16_spec  # noqa: E402  (pure numpy generator)


def _q(rng, n, robot):
    return rng.uniform(-np.pi, np.pi, (n, robot.ndof))


def _ref_scene(ab, sides, tfs):
    return ab.Scene([ab.Box(*s) for s in sides], [np.array(t) for t in tfs])


def _num_jac(fun, q, h=1e-5):
    f0 = fun(q)
    J = np.zeros((len(f0), len(q)))
    for c in range(len(q)):
        e = np.zeros(len(q))
        e[c] = h
        J[:, c] = (fun(q + e) - fun(q - e)) / (2 * h)
    return J


def _cc10k_chunk(args):
    """Worker: verdicts of the verbatim reference for one chunk of configurations."""
    q, = args
    ab = ref_shim.load_reference()
    kuka = ab.Kuka()
    s16_sides, s16_tfs = scene16_spec(0)
    s16 = _ref_scene(ab, s16_sides, s16_tfs)
    readme = _ref_scene(ab, np.array([[2, 2, 0.1], [0.2, 0.2, 1.5]]),
                        np.array([translation(0, 0, -0.2), translation(0, 0.5, 0.55)]))
    return (np.array([kuka.is_in_collision(x, s16) for x in q]),
            np.array([kuka.is_in_collision(x, readme) for x in q]),
            np.array([kuka.is_in_self_collision(x) for x in q]))


def make_cc10k(n=10_000, seed=20241020):
    """10 000 configurations through the verbatim Robot.is_in_collision (robot.py:244-273)."""
    import multiprocessing as mp

    q = np.random.default_rng(seed).uniform(-np.pi, np.pi, (n, 6))
    cores = min(8, os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(cores) as pool:
        parts = pool.map(_cc10k_chunk, [(c,) for c in np.array_split(q, cores * 4)])
    np.savez_compressed(
        os.path.join(GOLDEN, "cc_kuka_10k.npz"), seed=seed, q=q,
        hit_scene16=np.packbits(np.concatenate([p[0] for p in parts])),
        hit_readme=np.packbits(np.concatenate([p[1] for p in parts])),
        hit_self=np.packbits(np.concatenate([p[2] for p in parts])),
    )


def make_polyhedron(n=200, seed=7):
    ab = ref_shim.load_reference()
    rng = np.random.default_rng(seed)
    sides = rng.uniform(0.01, 2.0, (n, 3))
    tfs = np.array([pose_z(rng.uniform(-3, 3), *rng.uniform(-2, 2, 3)) @ pose_x(rng.uniform(-3, 3), 0, 0, 0)
                    @ pose_z(rng.uniform(-3, 3), 0, 0, 0) for _ in range(n)])
    A, b = [], []
    for s, tf in zip(sides, tfs):
        poly = ab.Box(*s).get_polyhedron(tf)
        A.append(np.array(poly.A))
        b.append(np.array(poly.b))
    np.savez(os.path.join(GOLDEN, "polyhedron.npz"), sides=sides, tf=tfs, A=np.array(A), b=np.array(b))


def main():
    ab = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    if "cc10k" in sys.argv[1:] or "polyhedron" in sys.argv[1:]:
        if "cc10k" in sys.argv[1:]:
            make_cc10k()
        if "polyhedron" in sys.argv[1:]:
            make_polyhedron()
        return
    rng = np.random.default_rng(20240601)

    # ---------------- FK ----------------
    for name in ["Kuka", "KukaOnRail", "PlanarArm", "SphericalArm", "SphericalWrist",
                 "AnthropomorphicArm", "Arm2"]:
        r = getattr(ab, name)()
        q = _q(rng, 64, r)
        np.savez(
            os.path.join(GOLDEN, f"fk_{name}.npz"),
            q=q,
            fk=np.array([r.fk(x) for x in q]),
            fk_all_links=np.array([r.fk_all_links(x) for x in q]),
            fk_rpy=np.array([r.fk_rpy(x) for x in q]),
        )
    r = ab.Kuka()
    tf_base = pose_z(0.7, 0.3, -0.2, 0.1) @ pose_x(-0.4, 0, 0, 0)
    tf_tip = pose_x(0.5, 0.05, -0.02, 0.2)
    r.tf_base = tf_base
    r.tf_tool_tip = tf_tip
    q = _q(rng, 64, r)
    np.savez(
        os.path.join(GOLDEN, "fk_Kuka_base_tip.npz"),
        q=q, tf_base=tf_base, tf_tool_tip=tf_tip,
        fk=np.array([r.fk(x) for x in q]),
        fk_all_links=np.array([r.fk_all_links(x) for x in q]),
        fk_rpy=np.array([r.fk_rpy(x) for x in q]),
    )

    # ---------------- collision ----------------
    out = {}
    kuka = ab.Kuka()
    readme_sides = np.array([[2, 2, 0.1], [0.2, 0.2, 1.5]])
    readme_tfs = np.array([translation(0, 0, -0.2), translation(0, 0.5, 0.55)])
    readme = _ref_scene(ab, readme_sides, readme_tfs)
    q_path = np.linspace([0.5, 1.5, -0.3, 0, 0, 0], [2.5, 1.5, 0.3, 0, 0, 0], 10)
    out["readme_q"] = q_path
    out["readme_hit"] = np.array([kuka.is_in_collision(x, readme) for x in q_path])
    # tests/test_robot.py
    t_sides = np.array([[0.2, 0.3, 0.5], [0.1, 0.3, 0.1]])
    out["test_robot_sides"] = t_sides
    out["test_robot_tfs1"] = np.array([pose_x(0, 0.75, 0, 0.5), pose_x(0, 0.75, 0.5, 0.5)])
    out["test_robot_tfs2"] = np.array([pose_x(0, 0.3, -0.7, 0.5), pose_x(0, 0.75, 0.5, 0.5)])
    q0 = np.array([0, np.pi / 2, 0, 0, 0, 0])
    out["test_robot_q0"] = q0
    out["test_robot_hit1"] = kuka.is_in_collision(q0, _ref_scene(ab, t_sides, out["test_robot_tfs1"]))
    out["test_robot_hit2"] = kuka.is_in_collision(q0, _ref_scene(ab, t_sides, out["test_robot_tfs2"]))
    q_self = np.array([[0, np.pi / 2, 0, 0, 0, 0], [0, 1.5, -1.3, 0, -1.5, 0]])
    out["self_q"] = q_self
    out["self_hit"] = np.array([kuka.is_in_self_collision(x) for x in q_self])
    # random configurations
    q = _q(rng, 1500, kuka)
    out["rand_q"] = q
    out["rand_hit_readme"] = np.array([kuka.is_in_collision(x, readme) for x in q])
    out["rand_hit_self"] = np.array([kuka.is_in_self_collision(x) for x in q])
    s16_sides, s16_tfs = scene16_spec(0)
    s16 = _ref_scene(ab, s16_sides, s16_tfs)
    out["rand_hit_scene16"] = np.array([kuka.is_in_collision(x, s16) for x in q])
    kuka.do_check_self_collision = False
    out["rand_hit_scene16_noself"] = np.array([kuka.is_in_collision(x, s16) for x in q])
    kuka.do_check_self_collision = True
    # Kuka + box-only welding torch + base geometry + moved base
    k2 = ab.Kuka()
    k2.tool = ab.torch2
    k2.tf_base = pose_z(0.3, 0.1, 0.05, 0.02)
    base_sides = np.array([[0.4, 0.4, 0.2]])
    base_tfs = np.array([translation(0, 0, -0.1)])
    k2.geometry_base = _ref_scene(ab, base_sides, base_tfs)
    q2 = _q(rng, 600, k2)
    out["tool_sides"] = np.array([[s.dx, s.dy, s.dz] for s in ab.torch2.s])
    out["tool_tfs"] = np.array(ab.torch2.tf_s)
    out["tool_tip"] = np.array(ab.torch2.tf_tool_tip)
    out["tool_base_tf"] = k2.tf_base
    out["tool_base_sides"] = base_sides
    out["tool_base_tfs"] = base_tfs
    out["tool_q"] = q2
    out["tool_hit_scene16"] = np.array([k2.is_in_collision(x, s16) for x in q2])
    out["tool_hit_self"] = np.array([k2.is_in_self_collision(x) for x in q2])
    out["tool_fk"] = np.array([k2.fk(x) for x in q2[:32]])
    # KukaOnRail (prismatic first joint)
    rail = ab.KukaOnRail()
    q3 = _q(rng, 400, rail)
    q3[:, 0] = rng.uniform(-1.0, 1.0, len(q3))
    out["rail_q"] = q3
    out["rail_hit_scene16"] = np.array([rail.is_in_collision(x, s16) for x in q3])
    np.savez(os.path.join(GOLDEN, "cc_kuka.npz"), **out)

    # ---------------- IK ----------------
    kuka = ab.Kuka()
    out = {}
    q = np.vstack([[0.1, 0.2, 0.3, 0.4, 0.5, 0.6], _q(rng, 400, kuka)])
    T = np.array([kuka.fk(x) for x in q])
    sol = np.full((len(q), 8, 6), np.nan)
    cnt = np.zeros(len(q), dtype=np.int64)
    for i, Ti in enumerate(T):
        res = kuka.ik(Ti)
        if res.success:
            cnt[i] = len(res.solutions)
            sol[i, : cnt[i]] = np.array(res.solutions)
    out.update(q=q, T=T, sol=sol, count=cnt)
    # with base + tool tip
    kb = ab.Kuka()
    kb.tf_base = tf_base
    kb.tf_tool_tip = tf_tip
    qb = _q(rng, 200, kb)
    Tb = np.array([kb.fk(x) for x in qb])
    solb = np.full((len(qb), 8, 6), np.nan)
    cntb = np.zeros(len(qb), dtype=np.int64)
    for i, Ti in enumerate(Tb):
        res = kb.ik(Ti)
        if res.success:
            cntb[i] = len(res.solutions)
            solb[i, : cntb[i]] = np.array(res.solutions)
    out.update(bt_base=tf_base, bt_tip=tf_tip, bt_T=Tb, bt_sol=solb, bt_count=cntb)
    # unreachable poses
    Tfar = T[:8].copy()
    Tfar[:, :3, 3] *= 3.0
    out["far_T"] = Tfar
    out["far_success"] = np.array([kuka.ik(Ti).success for Ti in Tfar])
    np.savez(os.path.join(GOLDEN, "ik_kuka.npz"), **out)

    out = {}
    arm = ab.Arm2(a1=0.18, a2=0.6, a3=0.62)
    p = rng.uniform(-1.2, 1.2, (300, 3))
    sols = np.full((len(p), 4, 3), np.nan)
    cnt = np.zeros(len(p), dtype=np.int64)
    for i, pi in enumerate(p):
        T = np.eye(4)
        T[:3, 3] = pi
        res = arm.ik(T)
        if res.success:
            cnt[i] = len(res.solutions)
            sols[i, : cnt[i]] = np.array(res.solutions)
    out.update(arm2_p=p, arm2_sol=sols, arm2_count=cnt)
    anth = ab.AnthropomorphicArm()
    p = rng.uniform(-1.5, 1.5, (300, 3))
    sols = np.full((len(p), 4, 3), np.nan)
    ok = np.zeros(len(p), dtype=bool)
    for i, pi in enumerate(p):
        T = np.eye(4)
        T[:3, 3] = pi
        res = anth.ik(T)
        ok[i] = res.success
        if res.success:
            sols[i] = np.array(res.solutions)
    out.update(anth_p=p, anth_sol=sols, anth_ok=ok)
    wrist = ab.SphericalWrist()
    qw = rng.uniform(-np.pi, np.pi, (200, 3))
    Rw = np.array([wrist.fk(x)[:3, :3] for x in qw])
    Rb = pose_z(0.4, 0, 0, 0) @ pose_x(0.9, 0, 0, 0)
    wrist.tf_base = Rb
    out.update(wrist_R=Rw, wrist_Rb=Rb[:3, :3],
               wrist_sol=np.array([wrist.ik(np.block([[R, np.zeros((3, 1))], [np.zeros((1, 3)), 1]])).solutions for R in Rw]))
    np.savez(os.path.join(GOLDEN, "ik_parts.npz"), **out)

    # ---------------- path sweep ----------------
    kuka = ab.Kuka()
    qs = _q(rng, 60, kuka) * 0.5
    qg = qs + rng.uniform(-0.6, 0.6, qs.shape)
    qg[0] = qs[0]  # zero-length segment: checks nothing (robot.py:233-237)
    qg[1] = qs[1] + 1e-3  # one step: checks q_start only
    np.savez(
        os.path.join(GOLDEN, "path.npz"),
        q_start=qs, q_goal=qg,
        n_steps=np.array([len(kuka._linear_interpolation_path(a, b, 0.1)) for a, b in zip(qs, qg)]),
        hit_readme=np.array([kuka.is_path_in_collision_discrete(a, b, readme) for a, b in zip(qs, qg)]),
        hit_scene16=np.array([kuka.is_path_in_collision_discrete(a, b, s16, 0.05) for a, b in zip(qs, qg)]),
    )

    # ---------------- box-box ----------------
    out = {}
    I = np.eye(4)
    far = np.eye(4)
    far[0, 3] = 10.0
    cases = [  # (sides1, tf1, sides2, tf2)
        ((1, 1, 1), I, (1, 1, 2), I),
        ((1, 1, 1), I, (1, 2, 1), pose_z(np.pi / 4, 0.7, 0.7, 0)),
        ((1, 1, 1), pose_z(0, -1, -1, 0), (1, 1, 2), pose_z(np.pi / 4, -2, -2, 0)),
        ((1, 1, 1), far, (1, 1, 2), I),
        ((1, 1, 1), I, (1, 1, 2), far),
    ]
    out["known_sides1"] = np.array([c[0] for c in cases], float)
    out["known_tf1"] = np.array([c[1] for c in cases])
    out["known_sides2"] = np.array([c[2] for c in cases], float)
    out["known_tf2"] = np.array([c[3] for c in cases])
    out["known_hit"] = np.array(
        [bool(ab.Box(*c[0]).is_in_collision(c[1], ab.Box(*c[2]), c[3])) for c in cases]
    )
    n = 3000
    s1 = rng.uniform(0.05, 1.0, (n, 3))
    s2 = rng.uniform(0.05, 1.0, (n, 3))

    def rand_tf(scale):
        qq = rng.normal(size=4)
        w, x, y, z = qq / np.linalg.norm(qq)
        tf = np.eye(4)
        tf[:3, :3] = [
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
        ]
        tf[:3, 3] = rng.uniform(-scale, scale, 3)
        return tf

    tf1 = np.array([rand_tf(0.6) for _ in range(n)])
    tf2 = np.array([rand_tf(0.6) for _ in range(n)])
    out.update(rand_sides1=s1, rand_tf1=tf1, rand_sides2=s2, rand_tf2=tf2)
    out["rand_hit"] = np.array(
        [bool(ab.Box(*a).is_in_collision(t1, ab.Box(*b), t2)) for a, t1, b, t2 in zip(s1, tf1, s2, tf2)]
    )
    np.savez(os.path.join(GOLDEN, "boxbox.npz"), **out)

    # ---------------- Jacobians (finite differences of the verbatim fk) ----------
    out = {}
    for name in ["Kuka", "PlanarArm", "KukaOnRail"]:
        r = getattr(ab, name)()
        q = rng.uniform(0.1, 1.0, (12, r.ndof))
        out[f"{name}_q"] = q
        out[f"{name}_jpos"] = np.array([_num_jac(lambda x: r.fk(x)[:3, 3].copy(), x) for x in q])
        out[f"{name}_jrpy"] = np.array([_num_jac(lambda x: r.fk_rpy(x), x) for x in q])
    np.savez(os.path.join(GOLDEN, "jac.npz"), **out)
    make_cc10k()
    make_polyhedron()
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
