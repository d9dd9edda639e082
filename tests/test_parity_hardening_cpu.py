"""CPU parity hardening:

* every restatement that a GPU result or a timed CPU arm is compared with - the vectorised
  numpy oracle, the plain-C oracle (``oracle_fk_cc``) and the per-configuration port the
  benchmark times as the reference arm (``oracle/scalar_port.py``) - is pinned to the verdicts
  of the VERBATIM reference (``tests/golden/cc_kuka*.npz``, 10 000 configurations);
* the restated FCL ``boxBox2`` predicate is checked against an FCL-independent ground truth:
  LP feasibility of the reference's own polyhedron description of a box
  (``Box.get_polyhedron``, reference ``src/acrobotics/shapes.py:163-172``) on > 10 000 random,
  structured and near-contact pairs.  Outside a 1e-5 m band around contact the verdicts must
  be identical; the band population is printed.
"""
import numpy as np
import pytest
from conftest import golden
from helpers import kuka_with_tool, spec

import acrobotics_b200 as ab
from acrobotics_b200.synthetic import readme_scene, scene16
from oracle import c_oracle, fcl_boxbox, lp_boxbox, scalar_port
from oracle import np_oracle as no

BAND = 1.0e-5  # metres: |LP depth| below this is "near contact" and only counted


def _bits(a, n):
    return np.unpackbits(a)[:n].astype(bool)


# ---------------------------------------------------------------------------
# restatements vs the verbatim reference
# ---------------------------------------------------------------------------
def test_c_oracle_and_numpy_oracle_match_verbatim_10k():
    g = golden("cc_kuka_10k.npz")
    q = g["q"]
    n = len(q)
    assert n >= 10_000
    s = spec(ab.Kuka())
    want16 = _bits(g["hit_scene16"], n)
    want_readme = _bits(g["hit_readme"], n)
    want_self = _bits(g["hit_self"], n)
    assert 0.5 < want16.mean() < 0.95 and 0.05 < want_self.mean() < 0.3
    # plain-C restatement, with the decision quantity
    hit, gmin = c_oracle.fk_cc(s, scene16(), q, want_g=True)
    assert (hit == want16).all()
    assert (np.abs(gmin) > 1e-6).all(), "a golden configuration sits inside the 1e-6 m margin"
    assert (c_oracle.fk_cc(s, readme_scene(), q) == want_readme).all()
    assert (c_oracle.fk_cc(s, None, q) == want_self).all()
    # vectorised numpy restatement
    assert (no.is_in_collision(s, scene16(), q) == want16).all()
    assert (no.is_in_collision(s, readme_scene(), q) == want_readme).all()
    assert (no.is_in_self_collision(s, q) == want_self).all()


def test_scalar_port_matches_verbatim():
    """The timed reference arm of bench.py returns the reference's verdicts."""
    g = golden("cc_kuka_10k.npz")
    m = 3000
    q = g["q"][:m]
    s = spec(ab.Kuka())
    for scene, key in ((scene16(), "hit_scene16"), (readme_scene(), "hit_readme")):
        got = scalar_port.run_chunk((s, no.scene_arrays(scene), q))
        assert (got == _bits(g[key], len(g["q"]))[:m]).all()
    # FK of the port against the verbatim golden poses
    f = golden("fk_Kuka.npz")
    bot = scalar_port.ScalarRobot(s)
    assert max(np.abs(bot.fk(x) - T).max() for x, T in zip(f["q"], f["fk"])) < 1e-14
    # tool + base geometry + moved base, prismatic rail
    c = golden("cc_kuka.npz")
    k = kuka_with_tool(c)
    got = scalar_port.run_chunk((spec(k), no.scene_arrays(scene16()), c["tool_q"]))
    assert (got == c["tool_hit_scene16"]).all()
    got = scalar_port.run_chunk((spec(ab.KukaOnRail()), no.scene_arrays(scene16()), c["rail_q"]))
    assert (got == c["rail_hit_scene16"]).all()
    # the C oracle on the same variants
    assert (c_oracle.fk_cc(spec(k), scene16(), c["tool_q"]) == c["tool_hit_scene16"]).all()
    assert (c_oracle.fk_cc(spec(ab.KukaOnRail()), scene16(), c["rail_q"]) == c["rail_hit_scene16"]).all()


def test_native_box_box_matches_scalar_restatement():
    g = golden("boxbox.npz")
    fn = c_oracle.box_box_fn()
    for pre in ("known", "rand"):
        s1, t1, s2, t2 = (np.ascontiguousarray(g[f"{pre}_{k}"]) for k in ("sides1", "tf1", "sides2", "tf2"))
        got = np.array([fn(s1[i], t1[i], s2[i], t2[i]) for i in range(len(s1))], dtype=bool)
        assert (got == g[f"{pre}_hit"]).all()


# ---------------------------------------------------------------------------
# LP ground truth
# ---------------------------------------------------------------------------
def test_polyhedron_matches_reference():
    g = golden("polyhedron.npz")
    for s, tf, A, b in zip(g["sides"], g["tf"], g["A"], g["b"]):
        A2, b2 = lp_boxbox.polyhedron(s, tf)
        assert np.abs(A2 - A).max() < 1e-15 and np.abs(b2 - b).max() < 1e-15


def _rand_rot(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0] = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1)
    R[:, 1] = np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1)
    R[:, 2] = np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    return R


def _tf(R, t):
    T = np.tile(np.eye(4), (len(R), 1, 1))
    T[:, :3, :3] = R
    T[:, :3, 3] = t
    return T


def _rz(a):
    c, s = np.cos(a), np.sin(a)
    R = np.zeros((len(a), 3, 3))
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, -s, s, c, 1.0
    return R


def _check_family(name, s1, t1, s2, t2):
    """SAT restatement vs LP depth; returns (pairs, band population, mismatches outside)."""
    hit, gq = no.sat(0.5 * s1, t1[:, :3, :3], t1[:, :3, 3], 0.5 * s2, t2[:, :3, :3], t2[:, :3, 3])
    d = np.array([lp_boxbox.depth(s1[i], t1[i], s2[i], t2[i]) for i in range(len(s1))])
    band = np.abs(d) <= BAND
    bad = ((d > 0) != hit) & ~band
    print(f"[lp-boxbox] {name}: {len(s1)} pairs, hit rate {hit.mean():.3f}, "
          f"{int(band.sum())} within +-{BAND:g} m of contact, {int(bad.sum())} mismatches outside")
    assert not bad.any(), (name, np.nonzero(bad)[0][:5], d[bad][:5], gq[bad][:5])
    return len(s1), int(band.sum())


def test_boxbox_restatement_equals_geometric_intersection():
    rng = np.random.default_rng(11)
    total = near = 0
    # (1) 8 000 random oriented pairs (sizes and offsets of the robot/scene scale)
    n = 8000
    s1, s2 = rng.uniform(0.03, 1.0, (n, 3)), rng.uniform(0.03, 1.0, (n, 3))
    t1 = _tf(_rand_rot(rng, n), rng.uniform(-0.5, 0.5, (n, 3)))
    t2 = _tf(_rand_rot(rng, n), rng.uniform(-0.5, 0.5, (n, 3)))
    a, b = _check_family("random", s1, t1, s2, t2)
    total, near = total + a, near + b
    # (2) parallel edges: box 2 is box 1's frame turned about a common axis (degenerate
    #     edge x edge axes - where ODE's fudge factor acts), incl. exactly aligned boxes
    n = 1200
    R1 = _rand_rot(rng, n)
    ang = rng.uniform(-np.pi, np.pi, n)
    ang[:200] = 0.0
    ang[200:400] = rng.choice([np.pi / 2, np.pi, -np.pi / 2], 200)
    ang[400:600] = rng.uniform(-1e-7, 1e-7, 200)
    s1, s2 = rng.uniform(0.05, 0.8, (n, 3)), rng.uniform(0.05, 0.8, (n, 3))
    t1 = _tf(R1, rng.uniform(-0.4, 0.4, (n, 3)))
    t2 = _tf(R1 @ _rz(ang), rng.uniform(-0.4, 0.4, (n, 3)))
    a, b = _check_family("parallel-axis", s1, t1, s2, t2)
    total, near = total + a, near + b
    # (3) thin plates (1 mm) and long bars against ordinary boxes
    n = 1200
    s1 = rng.uniform(0.05, 1.0, (n, 3))
    s1[:600, 2] = 1e-3
    s1[600:, :2] = rng.uniform(0.005, 0.02, (600, 2))
    s1[600:, 2] = rng.uniform(1.0, 3.0, 600)
    s2 = rng.uniform(0.05, 0.6, (n, 3))
    t1 = _tf(_rand_rot(rng, n), rng.uniform(-0.3, 0.3, (n, 3)))
    t2 = _tf(_rand_rot(rng, n), rng.uniform(-0.3, 0.3, (n, 3)))
    a, b = _check_family("plates+bars", s1, t1, s2, t2)
    total, near = total + a, near + b
    assert total >= 10_000
    print(f"[lp-boxbox] total {total} pairs, {near} inside the contact band")


def test_boxbox_near_contact_pairs():
    """Pairs pushed to a prescribed LP depth (bisection along the centre line): at +-1e-4 m and
    +-2e-5 m the restated boxBox2 must still agree with geometry; inside the band the
    disagreements are only counted."""
    rng = np.random.default_rng(5)
    n = 80
    s1, s2 = rng.uniform(0.05, 0.6, (n, 3)), rng.uniform(0.05, 0.6, (n, 3))
    R1, R2 = _rand_rot(rng, n), _rand_rot(rng, n)
    R2[:20] = R1[:20] @ _rz(rng.uniform(-np.pi, np.pi, 20))  # a quarter with parallel edges
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    counted = wrong_in_band = 0
    for target in (1e-4, -1e-4, 2e-5, -2e-5, 2e-6, -2e-6):
        for i in range(n):
            tf1 = _tf(R1[i:i + 1], np.zeros((1, 3)))[0]

            def depth_at(tau):
                tf2 = _tf(R2[i:i + 1], (tau * u[i])[None])[0]
                return lp_boxbox.depth(s1[i], tf1, s2[i], tf2), tf2

            lo, hi = 0.0, 3.0  # depth(lo) > 0 (concentric), depth(hi) < 0 (far apart)
            for _ in range(24):
                mid = 0.5 * (lo + hi)
                if depth_at(mid)[0] > target:
                    lo = mid
                else:
                    hi = mid
            d, tf2 = depth_at(0.5 * (lo + hi))
            hit, g = fcl_boxbox.box_box(s1[i], tf1[:3, :3], tf1[:3, 3], s2[i], tf2[:3, :3], tf2[:3, 3])
            if abs(d) > BAND:
                assert hit == (d > 0), (target, i, d, g)
                counted += 1
            else:
                wrong_in_band += int(hit != (d > 0))
    print(f"[lp-boxbox] near contact: {counted} pairs outside the band agree, "
          f"{wrong_in_band} of {2 * n} pairs at +-2e-6 m differ (fudge factor; reported only)")
    assert counted >= 4 * n - 8


@pytest.mark.parametrize("seed", [0, 1])
def test_boxbox_robot_scene_pairs(seed):
    """The box pairs the benchmark actually evaluates: Kuka link boxes at random configurations
    against the boxes of the 16-box scene."""
    rng = np.random.default_rng(100 + seed)
    s = spec(ab.Kuka())
    sh, sR, sT = no.scene_arrays(scene16())
    q = rng.uniform(-np.pi, np.pi, (40, 6))
    frames = no.fk_all_links(s, q)
    s1, t1, s2, t2 = [], [], [], []
    for c in range(len(q)):
        for b in s.boxes:
            W = frames[c, b.group] @ b.tf
            j = rng.integers(0, len(sh), 3)
            for k in j:
                T2 = np.eye(4)
                T2[:3, :3], T2[:3, 3] = sR[k], sT[k]
                s1.append(2 * b.half)
                t1.append(W)
                s2.append(2 * sh[k])
                t2.append(T2)
    _check_family(f"kuka x scene16 [{seed}]", np.array(s1), np.array(t1), np.array(s2), np.array(t2))
