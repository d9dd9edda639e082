"""The oracle is pinned here: oracle/np_oracle.py (and the scalar SAT) against the
golden vectors generated from the VERBATIM reference (oracle/make_golden.py) and
against the known answers the reference's own tests / README hold for this path.
CPU only."""
import numpy as np
import pytest
from conftest import golden
from helpers import ROBOTS, kuka_with_tool, product_scene, spec

import acrobotics_b200 as ab
from acrobotics_b200.synthetic import readme_path, readme_scene, scene16
from oracle import fcl_boxbox, np_oracle as no


@pytest.mark.parametrize("name", ROBOTS)
def test_fk_matches_reference(name):
    g = golden(f"fk_{name}.npz")
    s = spec(getattr(ab, name)())
    assert np.abs(no.fk(s, g["q"]) - g["fk"]).max() < 1e-14
    assert np.abs(no.fk_all_links(s, g["q"]) - g["fk_all_links"]).max() < 1e-14
    assert np.abs(no.fk_rpy(s, g["q"]) - g["fk_rpy"]).max() < 1e-13


def test_fk_base_and_tool_tip():
    g = golden("fk_Kuka_base_tip.npz")
    r = ab.Kuka()
    r.tf_base, r.tf_tool_tip = g["tf_base"], g["tf_tool_tip"]
    s = spec(r)
    assert np.abs(no.fk(s, g["q"]) - g["fk"]).max() < 1e-14
    assert np.abs(no.fk_all_links(s, g["q"]) - g["fk_all_links"]).max() < 1e-14


def test_readme_collision_path():
    # README.md:94-96
    expected = [False, False, False, False, True, True, True, True, False, False]
    hit = no.is_in_collision(spec(ab.Kuka()), readme_scene(), readme_path())
    assert hit.tolist() == expected
    assert golden("cc_kuka.npz")["readme_hit"].tolist() == expected


def test_reference_test_robot_cases():
    # tests/test_robot.py:15-41
    g = golden("cc_kuka.npz")
    s = spec(ab.Kuka())
    assert no.is_in_self_collision(s, g["self_q"]).tolist() == [False, True]
    sc1 = product_scene(g["test_robot_sides"], g["test_robot_tfs1"])
    sc2 = product_scene(g["test_robot_sides"], g["test_robot_tfs2"])
    assert bool(no.is_in_collision(s, sc1, g["test_robot_q0"])[0]) is True
    assert bool(no.is_in_collision(s, sc2, g["test_robot_q0"])[0]) is False


def test_collision_random_configs():
    g = golden("cc_kuka.npz")
    k = ab.Kuka()
    s = spec(k)
    q = g["rand_q"]
    assert (no.is_in_collision(s, readme_scene(), q) == g["rand_hit_readme"]).all()
    assert (no.is_in_self_collision(s, q) == g["rand_hit_self"]).all()
    assert (no.is_in_collision(s, scene16(), q) == g["rand_hit_scene16"]).all()
    k.do_check_self_collision = False
    assert (no.is_in_collision(spec(k), scene16(), q) == g["rand_hit_scene16_noself"]).all()
    assert 0.5 < g["rand_hit_scene16"].mean() < 0.95


def test_collision_tool_base_rail():
    g = golden("cc_kuka.npz")
    k = kuka_with_tool(g)
    s = spec(k)
    assert np.abs(no.fk(s, g["tool_q"][:32]) - g["tool_fk"]).max() < 1e-14
    assert (no.is_in_collision(s, scene16(), g["tool_q"]) == g["tool_hit_scene16"]).all()
    assert (no.is_in_self_collision(s, g["tool_q"]) == g["tool_hit_self"]).all()
    rail = spec(ab.KukaOnRail())
    assert (no.is_in_collision(rail, scene16(), g["rail_q"]) == g["rail_hit_scene16"]).all()


def test_boxbox_known_answers_and_random():
    g = golden("boxbox.npz")
    # tests/test_shapes.py:128-142 and tests/test_geometry.py:42-59
    assert g["known_hit"].tolist() == [True, True, False, False, False]
    for pre in ("known", "rand"):
        s1, t1, s2, t2 = (g[f"{pre}_{k}"] for k in ("sides1", "tf1", "sides2", "tf2"))
        hit, _ = no.sat(0.5 * s1, t1[:, :3, :3], t1[:, :3, 3], 0.5 * s2, t2[:, :3, :3], t2[:, :3, 3])
        assert (hit == g[f"{pre}_hit"]).all()
    # scalar and vectorised restatements agree, including the decision quantity
    for i in range(200):
        c, gq = fcl_boxbox.box_box(s1[i], t1[i, :3, :3], t1[i, :3, 3], s2[i], t2[i, :3, :3], t2[i, :3, 3])
        hv, gv = no.sat(0.5 * s1[i], t1[i, :3, :3], t1[i, :3, 3], 0.5 * s2[i], t2[i, :3, :3], t2[i, :3, 3])
        assert c == bool(hv) and abs(gq - float(gv)) < 1e-12


def test_scene_soup_cases():
    # tests/test_geometry.py:42-59 through the oracle's pairwise test
    g = golden("boxbox.npz")
    assert g["known_hit"][3:].tolist() == [False, False]


def test_ik_kuka_reference_solutions():
    g = golden("ik_kuka.npz")
    s = spec(ab.Kuka())
    sol, valid = no.ik_kuka(s, g["T"])
    assert (valid.sum(1) == g["count"]).all()
    for i in range(len(sol)):
        assert np.abs(sol[i][valid[i]] - g["sol"][i, : g["count"][i]]).max() < 1e-12
    # README.md:57-63
    readme = np.array(
        [
            [0.1, -1.0949727, 2.84159265, 2.87778828, 0.79803563, -1.99992985],
            [0.1, -1.0949727, 2.84159265, -0.26380438, -0.79803563, 1.1416628],
            [0.1, 0.2, 0.3, 0.4, 0.5, 0.6],
            [0.1, 0.2, 0.3, -2.74159265, -0.5, -2.54159265],
        ]
    )
    assert np.abs(sol[0][valid[0]] - readme).max() < 1e-7
    k = ab.Kuka()
    k.tf_base, k.tf_tool_tip = g["bt_base"], g["bt_tip"]
    sol, valid = no.ik_kuka(spec(k), g["bt_T"])
    assert (valid.sum(1) == g["bt_count"]).all()
    for i in range(len(sol)):
        assert np.abs(sol[i][valid[i]] - g["bt_sol"][i, : g["bt_count"][i]]).max() < 1e-12
    _, valid = no.ik_kuka(s, g["far_T"])
    assert (valid.any(1) == g["far_success"]).all()


def test_ik_round_trip():
    # tests/test_inverse_kinematics.py:107-143: fk(ik(T)) == T to 5 decimals
    s = spec(ab.Kuka())
    q = np.random.default_rng(3).uniform(-np.pi, np.pi, (300, 6))
    T = no.fk(s, q)
    sol, valid = no.ik_kuka(s, T)
    assert valid.any(1).all()
    for k in range(8):
        m = valid[:, k]
        assert np.abs(no.fk(s, sol[m, k]) - T[m]).max() < 1e-5


def test_ik_parts():
    g = golden("ik_parts.npz")
    sol, valid = no.ik_arm2(g["arm2_p"], 0.18, 0.6, 0.62)
    assert (valid.sum(1) == g["arm2_count"]).all()
    assert 0 < (g["arm2_count"] == 0).sum() < len(sol)  # some targets are out of reach
    for i in np.nonzero(g["arm2_count"])[0]:
        assert np.abs(sol[i][valid[i]] - g["arm2_sol"][i, : g["arm2_count"][i]]).max() < 1e-13
    sol, ok = no.ik_anthropomorphic(g["anth_p"], 1.0, 1.0)
    assert (ok == g["anth_ok"]).all()
    assert np.abs(sol[ok] - g["anth_sol"][ok]).max() < 1e-13
    sol = no.ik_spherical_wrist(g["wrist_R"], np.broadcast_to(g["wrist_Rb"], g["wrist_R"].shape))
    assert np.abs(sol - g["wrist_sol"]).max() < 1e-13


def test_path_sweep():
    g = golden("path.npz")
    s = spec(ab.Kuka())
    n = [len(no.interpolation_path(a, b, 0.1)) for a, b in zip(g["q_start"], g["q_goal"])]
    assert n == g["n_steps"].tolist()
    assert n[0] == 0 and n[1] == 1
    hit = no.is_path_in_collision_discrete(s, readme_scene(), g["q_start"], g["q_goal"], 0.1)
    assert (hit == g["hit_readme"]).all()
    hit = no.is_path_in_collision_discrete(s, scene16(), g["q_start"], g["q_goal"], 0.05)
    assert (hit == g["hit_scene16"]).all()


@pytest.mark.parametrize("name", ["Kuka", "PlanarArm", "KukaOnRail"])
def test_jacobians_against_finite_differences(name):
    # the method of reference tests/test_jacobians.py:9-18, applied to the verbatim fk
    g = golden("jac.npz")
    s = spec(getattr(ab, name)())
    q = g[f"{name}_q"]
    assert np.abs(no.jacobian_position(s, q) - g[f"{name}_jpos"]).max() < 1e-8
    J = no.jacobian_rpy(s, q)
    assert np.abs(J[:, :3] - g[f"{name}_jpos"]).max() < 1e-8
    assert np.abs(J[:, 3:] - g[f"{name}_jrpy"][:, 3:]).max() < 1e-6
