"""Host logic of the path package against the reference's own test expectations
(tests/test_path.py, tests/test_path_util.py of the reference)."""
import numpy as np

from acrobotics_b200.path import (NoTolerance, SymmetricTolerance, Tolerance, TolEulerPt,
                                  TolPositionPt, create_grid, rpy_to_rot_mat)


def test_create_grid_docstring_example():
    # src/acrobotics/path/util.py:44-52
    g = create_grid([np.array([0, 1]), np.array([1, 2, 3]), np.array([99])])
    assert g.tolist() == [[0, 1, 99], [1, 1, 99], [0, 2, 99], [1, 2, 99], [0, 3, 99], [1, 3, 99]]


def test_tolerances():
    t = Tolerance(0.2, 1.0, 3)
    assert np.allclose(t.discretize(), [0.2, 0.6, 1.0]) and t.has_tolerance
    assert np.allclose(SymmetricTolerance(0.1, 3).discretize(), [-0.1, 0.0, 0.1])
    n = NoTolerance()
    assert not n.has_tolerance and n.discretize().tolist() == [0]
    t.reduce_tolerance(2.0, 0.5)
    assert t.lower <= t.upper


def test_position_point_grid():
    # reference tests/test_path.py: TolPositionPt with a frame rotated 90 deg about z
    Rz = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    tol = [SymmetricTolerance(0.1, 3), Tolerance(0.2, 1.0, 3), NoTolerance()]
    pt = TolPositionPt([1, 2, 3], Rz, tol)
    g = pt.sample_grid_array()
    assert g.shape == (9, 4, 4)
    assert np.allclose(g[:, 2, 3], 3.0) and np.allclose(g[:, :3, :3], Rz)
    # local x tolerance acts along world y, local y tolerance along world -x
    assert g[:, 1, 3].min() >= 1.9 - 1e-12 and g[:, 1, 3].max() <= 2.1 + 1e-12
    assert g[:, 0, 3].min() >= 0.0 - 1e-12 and g[:, 0, 3].max() <= 0.8 + 1e-12
    assert pt.sample_dim == 2 and len(pt.sample_grid()) == 9


def test_euler_point_grid():
    # reference tests/test_path.py:133-142
    pt = TolEulerPt([1, 2, 3], np.eye(3), [NoTolerance()] * 3,
                    [NoTolerance(), NoTolerance(), SymmetricTolerance(np.pi, 3)])
    tf = pt.sample_grid_array()
    assert np.allclose(tf[0], tf[2]) and np.allclose(tf[1][:3, :3], np.eye(3))
    assert np.allclose(tf[:, :3, 3], [1, 2, 3])


def test_rpy_convention_inverts_fk_rpy_formula():
    # src/acrobotics/robot.py:71-80 extracts r_x, r_y from R = Rx Ry Rz
    rng = np.random.default_rng(0)
    a = rng.uniform(-1.2, 1.2, (50, 3))
    R = rpy_to_rot_mat(a)
    assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3), atol=1e-12)
    assert np.allclose(np.arctan2(-R[:, 1, 2], R[:, 2, 2]), a[:, 0])
    assert np.allclose(np.arctan2(R[:, 0, 2], np.hypot(R[:, 0, 0], R[:, 0, 1])), a[:, 1])
    assert np.allclose(np.arctan2(-R[:, 0, 1], R[:, 0, 0]), a[:, 2])


def test_path_grid_arrays_equal_per_point_grids():
    from acrobotics_b200.path import path_grid_arrays

    rng = np.random.default_rng(0)

    def rot(a):
        return rpy_to_rot_mat(np.array(a))

    path = []
    for k in range(7):
        path.append(TolEulerPt(rng.normal(size=3), rot(rng.normal(size=3)),
                               [SymmetricTolerance(0.1, 3), NoTolerance(), NoTolerance()],
                               [NoTolerance(), SymmetricTolerance(0.5, 4), SymmetricTolerance(np.pi, 5)]))
        path.append(TolPositionPt(rng.normal(size=3), rot(rng.normal(size=3)),
                                  [SymmetricTolerance(0.1, 3), Tolerance(0.2, 1.0, 2), NoTolerance()]))
    seen = set()
    for idxs, T in path_grid_arrays(path):
        for k, i in enumerate(idxs):
            assert np.allclose(T[k], path[i].sample_grid_array(), atol=1e-14)
            seen.add(i)
    assert seen == set(range(len(path)))


# ---------------------------------------------------------------------------
# the reference's tests/test_path.py, restated for this package (no GPU needed: the robots
# are the reference's own stubs with `ik` and `is_in_collision` only)
# ---------------------------------------------------------------------------
import pytest  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200.path import (QuaternionTolerance, SampleMethod, SamplingSetting,  # noqa: E402
                                  SearchStrategy, TolQuatPt, create_arc, create_line,
                                  rotation_matrix_to_rpy)
from acrobotics_b200.path.path_pt import _PathPt  # noqa: E402
from acrobotics_b200.path.sampling import Sampler  # noqa: E402

RZ90 = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
IK_RESULT = ab.IKResult(True, [np.ones(6), 2 * np.ones(6)])


class DummyRobot:
    """tests/test_path.py:25-35"""

    def __init__(self, is_colliding=False):
        self.is_colliding = is_colliding
        self.ndof = 6

    def is_in_collision(self, joint_position, scene=None):
        return self.is_colliding

    def ik(self, transform):
        return IK_RESULT


def test_calc_ik():
    js = _PathPt._calc_ik(DummyRobot(), [[], []])
    assert len(js) == 4
    for got, want in zip(js, IK_RESULT.solutions * 2):
        assert np.allclose(got, want)


def test_position_pt_reference_grids():
    # tests/test_path.py:60-85
    pt = TolPositionPt([0.5, 0.5, 1], np.eye(3),
                       [Tolerance(-0.5, 0.5, 2), Tolerance(-0.5, 0.5, 2), NoTolerance()])
    for pos, T in zip([[0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]], pt.sample_grid()):
        assert np.allclose(pos, T[:3, 3])
    pt = TolPositionPt([1, 2, 3], RZ90, [SymmetricTolerance(0.1, 3), NoTolerance(), NoTolerance()])
    assert np.allclose([tf[:3, 3] for tf in pt.sample_grid()], [[1, 1.9, 3], [1, 2, 3], [1, 2.1, 3]])


@pytest.mark.parametrize("method", list(SampleMethod))
def test_position_pt_sample_incremental(method):
    # tests/test_path.py:87-122
    pt = TolPositionPt([1, 2, 3], RZ90, [SymmetricTolerance(0.1, 3), NoTolerance(), NoTolerance()])
    g = pt.sample_incremental(10, method)
    assert len(g) == 10
    for tf in g:
        assert np.isclose(tf[0, 3], 1) and np.isclose(tf[2, 3], 3) and 1.9 <= tf[1, 3] <= 2.1
    pt = TolPositionPt([1, 2, 3], RZ90, [SymmetricTolerance(0.1, 3), Tolerance(0.2, 1.0, 3), NoTolerance()])
    for tf in pt.sample_incremental(10, method):
        assert np.isclose(tf[2, 3], 3) and 0 <= tf[0, 3] <= 3 and 1.9 <= tf[1, 3] <= 2.1


def test_rel_tolerance_deviation():
    # tests/test_path.py:124-136 and :158-176
    tol = [SymmetricTolerance(0.1, 3), Tolerance(0.2, 1.0, 3), NoTolerance()]
    tf = np.eye(4)
    tf[:3, :3] = RZ90
    tf[:3, 3] = [1.06, 1.91, 3]
    assert np.allclose(TolPositionPt([1, 2, 3], RZ90, tol).transform_to_rel_tolerance_deviation(tf),
                       [-0.09, -0.06, 0.0])
    rot_tol = [NoTolerance(), NoTolerance(), SymmetricTolerance(0.1, 3)]
    pt = TolEulerPt([1, 2, 3], RZ90, tol, rot_tol)
    a = np.pi / 2 - 0.05
    tf[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    p_rel = pt.transform_to_rel_tolerance_deviation(tf)
    assert p_rel.shape == (6,) and np.allclose(p_rel, [-0.09, -0.06, 0.0, 0.0, 0.0, -0.05])
    # batched form (used by calc_state_cost)
    assert np.allclose(pt.transform_to_rel_tolerance_deviation(np.stack([tf, tf])), [p_rel, p_rel])


def test_euler_pt_incremental():
    # tests/test_path.py:144-156
    pt = TolEulerPt([1, 2, 3], np.eye(3), [NoTolerance()] * 3,
                    [NoTolerance(), NoTolerance(), SymmetricTolerance(np.pi, 3)])
    for tf in pt.sample_incremental(10, SampleMethod.random_uniform):
        e = rotation_matrix_to_rpy(tf[:3, :3])
        assert np.isclose(e[0], 0) and np.isclose(e[1], 0) and -np.pi <= e[2] <= np.pi


def test_quat_pt():
    # tests/test_path.py:179-210
    pt = TolQuatPt([1, 2, 3], np.eye(3), 3 * [NoTolerance()], QuaternionTolerance(0.1))
    for tf in pt.sample_incremental(100, SampleMethod.random_uniform):
        angle = np.arccos(np.clip(0.5 * (np.trace(tf[:3, :3]) - 1), -1, 1))
        assert 0.5 * angle <= 0.1 + 1e-12 and np.allclose(tf[:3, 3], [1, 2, 3])
    assert np.allclose(pt.to_transform([1, 2, 3], np.eye(3)),
                       [[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]])
    with pytest.raises(NotImplementedError):
        pt.sample_grid()


def _settings(strategy):
    return SamplingSetting(strategy, 1, SampleMethod.random_uniform, 10, 10, 10, 2)


def test_to_joint_solutions_all_strategies():
    # tests/test_path.py:238-255: a robot that always collides
    pt = TolPositionPt([1, 2, 3], RZ90, [SymmetricTolerance(0.1, 3), NoTolerance(), NoTolerance()])
    for strategy in SearchStrategy:
        if strategy is not SearchStrategy.MIN_INCREMENTAL:
            assert len(pt.to_joint_solutions(DummyRobot(is_colliding=True), _settings(strategy))) == 0
        else:
            with pytest.raises(Exception, match="Maximum iterations reached in to_joint_solutions."):
                pt.to_joint_solutions(DummyRobot(is_colliding=True), _settings(strategy))
    # and one that never does: 3 grid poses x 2 solutions; 10 samples x 2; >= 10 for the search
    free = DummyRobot(is_colliding=False)
    assert pt.to_joint_solutions(free, _settings(SearchStrategy.GRID)).shape == (6, 6)
    assert pt.to_joint_solutions(free, _settings(SearchStrategy.INCREMENTAL)).shape == (20, 6)
    assert pt.to_joint_solutions(free, _settings(SearchStrategy.MIN_INCREMENTAL)).shape == (10, 6)


def test_sampler_halton_is_incremental_and_in_range():
    s = Sampler()
    a = s.sample(5, 3, SampleMethod.deterministic_uniform)
    b = s.sample(5, 3, SampleMethod.deterministic_uniform)
    both = Sampler().sample(10, 3, SampleMethod.deterministic_uniform)
    assert np.allclose(np.vstack([a, b]), both) and (both > 0).all() and (both < 1).all()
    assert np.allclose(both[:3, 0], [0.5, 0.25, 0.75]) and np.allclose(both[:2, 1], [1 / 3, 2 / 3])
    r = Sampler(seed=1).sample(1000, 2, SampleMethod.random_uniform)
    assert r.shape == (1000, 2) and 0 <= r.min() and r.max() < 1


def test_sampling_setting_asserts_like_the_reference():
    with pytest.raises(AssertionError):
        SamplingSetting(SearchStrategy.INCREMENTAL)
    with pytest.raises(AssertionError):
        SamplingSetting(SearchStrategy.MIN_INCREMENTAL, sample_method=SampleMethod.random_uniform)
    with pytest.raises(AssertionError):
        SamplingSetting(SearchStrategy.GRID, iterations=2)
    s = SamplingSetting(SearchStrategy.MIN_INCREMENTAL, sample_method=SampleMethod.random_uniform,
                        desired_num_samples=5, max_search_iters=7)
    assert s.step_size == 1 and s.iterations == 1


def test_factories():
    # reference tests/test_path_factory style: end points and spacing
    start = TolPositionPt([0.9, -0.2, 0.2], np.eye(3), [NoTolerance()] * 3)
    line = create_line(start, np.array([0.9, 0.2, 0.2]), 5)
    assert len(line) == 5 and np.allclose(line[-1].pos, [0.9, 0.2, 0.2]) and line[0] is start
    arc = create_arc(start, np.array([0.9, 0.0, 0.2]), np.array([0, 0, 1]), np.pi, 3)
    assert np.allclose(arc[-1].pos, [0.9, 0.2, 0.2]) and np.allclose(arc[1].pos, [1.1, 0.0, 0.2])
    assert np.allclose(arc[-1].rotation_matrix, np.diag([-1.0, -1.0, 1.0]))
    with pytest.raises(Exception):
        create_line(start, np.zeros(3), 1)
