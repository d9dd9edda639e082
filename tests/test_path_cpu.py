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
