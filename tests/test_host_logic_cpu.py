"""Host-side logic that needs no GPU: flattening of robots / scenes into the ABI
descriptors, the reference-named API surface, plain box geometry."""
import numpy as np
import pytest
from conftest import golden
from helpers import kuka_with_tool, spec

import acrobotics_b200 as ab
from acrobotics_b200 import _native
from acrobotics_b200.synthetic import scene16, scene16_spec
from acrobotics_b200.transforms import pack_3x4, pose_x, pose_z, rot_x, rot_y, rot_z, tf_inverse, translation


def test_reference_module_paths_and_names():
    from acrobotics_b200.geometry import Scene
    from acrobotics_b200.inverse_kinematics import anthropomorphic_arm, arm_2, spherical_wrist
    from acrobotics_b200.link import DHLink, JointType, Link
    from acrobotics_b200.robot import IKResult, JointLimit, Robot, Tool
    from acrobotics_b200.robot_examples import Kuka
    from acrobotics_b200.shapes import Box

    assert Kuka is ab.Kuka and Box is ab.Box and Scene is ab.Scene and Robot is ab.Robot
    assert JointType.revolute.value == "r" and JointType.prismatic.value == "p"
    assert DHLink(1, 2, 3, 4)._fields == ("a", "alpha", "d", "theta")
    assert JointLimit(-1, 1).upper == 1
    assert not hasattr(IKResult(False), "solutions")
    for m in (arm_2, spherical_wrist, anthropomorphic_arm):
        assert callable(m.ik)
    with pytest.raises(ValueError):
        Link(DHLink(0, 0, 0, 0), "x", None)
    assert issubclass(Tool, Scene)


def test_kuka_tables_match_reference():
    g = golden("fk_Kuka.npz")
    k = ab.Kuka()
    assert k.ndof == 6 and len(k.joint_limits) == 6
    assert np.isclose(k.joint_limits[1].lower, np.deg2rad(-180))
    assert np.isclose(k.estimate_max_extension(), 0.18 + 0.6 + 0.62 + 0.115)
    # band collision matrix (robot.py:185-186)
    cm = k.collision_matrix
    assert cm.dtype == bool and cm.sum() == 12 and cm[0, 3] and cm[3, 0] and not cm[0, 2]
    assert k.do_check_self_collision is True and k.tf_tool_tip is None
    assert g["q"].shape[1] == k.ndof


def test_descriptor_flattening():
    k = ab.Kuka()
    d = k._describe()
    assert d.ndof == 6 and d.n_boxes == 6 and d.n_self_pairs == 6 and d.check_self == 1
    pairs = sorted(tuple(d.self_pairs[i]) for i in range(6))
    assert pairs == [(0, 3), (0, 4), (0, 5), (1, 4), (1, 5), (2, 5)]
    assert d.cos_alpha[0] == float(np.cos(np.pi / 2)) != 0.0  # 6.1e-17, like the reference
    assert list(d.boxes[1].half) == [0.4, 0.1, 0.05]
    assert d.boxes[1].tf[3] == -0.3 and d.boxes[1].carrier == 1
    k.do_check_self_collision = False
    assert k._describe().check_self == 0 and k._describe(force_self=True).check_self == 1
    # tool + base geometry
    g = golden("cc_kuka.npz")
    kt = kuka_with_tool(g)
    d = kt._describe()
    assert d.n_boxes == 6 + 6 + 1 and d.has_tool_tip == 1
    carriers = [d.boxes[i].carrier for i in range(d.n_boxes)]
    assert carriers.count(_native.CARRIER_TOOL) == 6 and carriers.count(_native.CARRIER_BASE) == 1
    assert d.n_self_pairs == 6 + 5 * 6  # link pairs + tool boxes x links[:-1]
    assert np.allclose(list(d.tf_base), pack_3x4(g["tool_base_tf"]))
    # the oracle's flattening sees the same robot
    s = spec(kt)
    assert len(s.boxes) == 13 and len(s.self_pairs) == 6 + 5
    # Jacobian descriptor: as constructed (identity base, no tool), robot.py:194-195
    dj = kt._describe(kinematics_only=True, as_constructed=True)
    assert dj.has_tool_tip == 0 and list(dj.tf_base)[:4] == [1, 0, 0, 0] and dj.n_boxes == 0
    rail = ab.KukaOnRail()._describe()
    assert rail.ndof == 7 and list(rail.joint_type)[:2] == [1, 0] and rail.n_self_pairs == 10


def test_limits():
    k = ab.Kuka()
    assert k._is_in_limits([0] * 6) and not k._is_in_limits([3.0, 0, 0, 0, 0, 0])
    with pytest.raises(ValueError):
        ab.Robot([ab.Kuka().links[0]] * 9)._describe()
    many = ab.Scene([ab.Box(1, 1, 1)] * 65, [np.eye(4)] * 65)
    assert many.packed().shape == (65, 15)


def test_interpolation_path_matches_reference_semantics():
    g = golden("path.npz")
    k = ab.Kuka()
    n = [len(k._linear_interpolation_path(a, b, 0.1)) for a, b in zip(g["q_start"], g["q_goal"])]
    assert n == g["n_steps"].tolist()
    P = k._linear_interpolation_path([0, 0, 0, 0, 0, 0], [0.25, 0, 0, 0, 0, 0], 0.1)
    assert len(P) == 3 and np.allclose(P[1], [0.125, 0, 0, 0, 0, 0]) and P[-1][0] == 0.25


def test_box_geometry():
    # reference tests/test_shapes.py:30-126 / examples/geometry_nb.ipynb
    b = ab.Box(1, 2, 3)
    v = b.get_vertices(np.eye(4))
    assert v.shape == (8, 3) and np.allclose(v[0], [-0.5, 1, 1.5]) and np.allclose(v[7], [0.5, -1, -1.5])
    tf = pose_z(0.3, 0.1, 0.2, -0.3)
    A, bb = b.get_polyhedron(tf)
    Aa = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]) @ tf[:3, :3].T
    assert np.allclose(A, Aa) and np.allclose(bb, np.array([0.5, 0.5, 1, 1, 1.5, 1.5]) + Aa @ tf[:3, 3])
    e = b.get_edges(tf)
    assert e.shape == (12, 6) and b.num_edges == 12
    lengths = sorted(np.round(np.linalg.norm(e[:, :3] - e[:, 3:], axis=1), 9))
    assert lengths == [1.0] * 4 + [2.0] * 4 + [3.0] * 4
    sc = ab.Scene([ab.Box(1, 1, 1)], [np.eye(4)])
    sc.add([ab.Box(2, 2, 2)], [translation(1, 0, 0)])
    sc.translate(0, 0, 1)
    assert len(sc.shapes) == 2 and sc.tf_s[1][2, 3] == 1 and len(sc.get_polyhedrons()) == 2


def test_transform_helpers():
    assert np.allclose(rot_x(0.3) @ rot_x(-0.3), np.eye(3))
    assert np.allclose(rot_y(np.pi / 2) @ [0, 0, 1], [1, 0, 0])
    assert np.allclose(rot_z(np.pi / 2) @ [1, 0, 0], [0, 1, 0])
    T = pose_x(0.4, 1, 2, 3)
    assert np.allclose(T @ tf_inverse(T), np.eye(4)) and np.allclose(T[:3, 3], [1, 2, 3])
    assert np.allclose(T[1:3, 1:3], [[np.cos(0.4), -np.sin(0.4)], [np.sin(0.4), np.cos(0.4)]])


def test_scene16_is_deterministic():
    s1, t1 = scene16_spec(0)
    s2, t2 = scene16_spec(0)
    assert s1.shape == (16, 3) and np.array_equal(t1, t2) and np.array_equal(s1, s2)
    assert np.allclose(s1[0], [2, 2, 0.1]) and np.allclose(t1[0][:3, 3], [0, 0, -0.2])
    R = t1[:, :3, :3]
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3))
    assert (np.hypot(t1[1:, 0, 3], t1[1:, 1, 3]) >= 0.45).all()
    assert scene16().packed().shape == (16, 15)
