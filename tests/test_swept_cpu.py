"""Swept collision (reference ``Robot.is_path_in_collision``, robot.py:285-317): the numpy
oracle and the scalar FCL restatement against the verbatim-reference golden vectors
(``tests/golden/ccd.npz``, made by ``oracle/make_golden_ccd.py``)."""
import numpy as np
import pytest
from conftest import golden

from oracle import fcl_ccd, np_oracle as no


def _spec(g, tool=True, base=False, rail=False):
    """Flattened Kuka (+ torch2, + base geometry) without the CUDA library."""
    import acrobotics_b200 as ab

    k = ab.KukaOnRail() if rail else ab.Kuka()
    if tool:
        k.tool = ab.Tool([ab.Box(*s) for s in g["tool_sides"]], list(g["tool_tfs"]), g["tool_tip"])
    if base:
        k.tf_base = g["full_base_tf"]
        k.geometry_base = ab.Scene([ab.Box(*s) for s in g["full_base_sides"]], list(g["full_base_tfs"]))
    return no.spec_from_robot(k)


def _scene(sides, tfs):
    tfs = np.asarray(tfs, float)
    return 0.5 * np.asarray(sides, float), tfs[:, :3, :3].copy(), tfs[:, :3, 3].copy()


def _scene16():
    from acrobotics_b200.synthetic import scene16_spec

    return _scene(*scene16_spec(0))


@pytest.mark.parametrize("case", ["s3", "s16", "full", "rail"])
def test_numpy_oracle_equals_verbatim_reference(case):
    g = golden("ccd.npz")
    spec = _spec(g, tool=case != "rail", base=case == "full", rail=case == "rail")
    scene = _scene(g["s3_sides"], g["s3_tfs"]) if case in ("s3", "full") else _scene16()
    gm = no.swept_margins(spec, scene, g[case + "_q_start"], g[case + "_q_goal"])
    hit = ~(gm > 0)
    ref = g[case + "_hit"]
    robust = np.abs(gm) > 1e-9
    assert np.array_equal(hit[robust], ref[robust])
    assert robust.mean() > 0.995
    assert 0.2 < ref.mean() < 0.9  # both verdicts are exercised


def test_known_answers_and_sticky_result():
    """test_continuous_cc.py:35-83 asserts True for its three set-ups on ONE shared robot.
    The stateless FCL query gives True, True, False; python-fcl's accumulating result object
    (one per Shape in the reference) turns the third into True."""
    g = golden("ccd.npz")
    assert g["known_sticky"].tolist() == [True, True, True]
    assert g["known_stateless"].tolist() == [True, True, False]
    spec = _spec(g)
    for sticky, want in ((True, g["known_sticky"]), (False, g["known_stateless"])):
        emu = fcl_ccd.StickyRobotEmulation(spec, sticky=sticky)
        got = [
            emu.is_path_in_collision(g["known_q_start"][i], g["known_q_goal"][i],
                                     _scene(g["known_sides"][i], g["known_tfs"][i]))
            for i in range(3)
        ]
        assert got == want.tolist()
    # the vectorised oracle is the stateless query
    for i in range(3):
        gm = no.swept_margins(spec, _scene(g["known_sides"][i], g["known_tfs"][i]),
                              g["known_q_start"][i], g["known_q_goal"][i])
        assert bool(gm[0] <= 0) == bool(g["known_stateless"][i])
        assert abs(gm[0]) > 1e-3  # none of the three is a near-contact case


def test_interp_motion_restatements_agree():
    """Scalar (oracle/fcl_ccd.py) and vectorised (np_oracle.interp_motion) InterpMotion."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        B = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        A *= np.sign(np.linalg.det(A))
        B *= np.sign(np.linalg.det(B))
        ta, tb = rng.normal(size=3), rng.normal(size=3)
        m = fcl_ccd.InterpMotion(A, ta, B, tb)
        R, p = no.interp_motion(A[None], ta[None], B[None], tb[None])
        for i in range(10):
            Ri, pi = m.at(i / 9)
            assert np.abs(R[0, i] - Ri).max() < 1e-14 and np.abs(p[0, i] - pi).max() < 1e-14
        # end poses are reproduced, the rotation angle is the geodesic one
        assert np.abs(R[0, 0] - A).max() < 1e-12 and np.abs(R[0, -1] - B).max() < 1e-12
        assert 0.0 <= m.angular_vel <= np.pi + 1e-12
    # no rotation: axis (1, 0, 0), angle 0 (Eigen's convention)
    ang, ax = fcl_ccd.angle_axis(np.eye(3))
    assert ang == 0.0 and ax.tolist() == [1.0, 0.0, 0.0]
