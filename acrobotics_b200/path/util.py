"""Grid helper of the path samplers (reference ``src/acrobotics/path/util.py:23-57``) and the
roll-pitch-yaw convention the samplers use."""
import numpy as np


def create_grid(r):
    """All combinations of N sampled ranges -> (M, N); first range varies second-fastest
    exactly as ``numpy.meshgrid`` (indexing='xy') + flatten does in the reference."""
    grid = np.meshgrid(*r)
    grid = [grid[i].flatten() for i in range(len(r))]
    return np.array(grid).T


def rpy_to_rot_mat(rxyz):
    """(..., 3) angles -> (..., 3, 3).

    ``acrolib.geometry.rpy_to_rot_mat`` is not vendored in the reference tree (unpinned
    dependency); the convention used here, R = Rx(r_x) Ry(r_y) Rz(r_z), is the one the
    reference's own ``fk_rpy`` inverts (``src/acrobotics/robot.py:71-80``) and satisfies the
    reference's tests on it (``tests/test_path.py:133-156``).  Parity of this helper with
    acrolib is otherwise unpinned; the GPU pipeline takes 4x4 transforms and does not depend on
    it."""
    rxyz = np.asarray(rxyz, dtype=np.float64)
    cx, cy, cz = np.cos(rxyz[..., 0]), np.cos(rxyz[..., 1]), np.cos(rxyz[..., 2])
    sx, sy, sz = np.sin(rxyz[..., 0]), np.sin(rxyz[..., 1]), np.sin(rxyz[..., 2])
    R = np.empty(rxyz.shape[:-1] + (3, 3))
    R[..., 0, 0] = cy * cz
    R[..., 0, 1] = -cy * sz
    R[..., 0, 2] = sy
    R[..., 1, 0] = cx * sz + sx * sy * cz
    R[..., 1, 1] = cx * cz - sx * sy * sz
    R[..., 1, 2] = -sx * cy
    R[..., 2, 0] = sx * sz - cx * sy * cz
    R[..., 2, 1] = sx * cz + cx * sy * sz
    R[..., 2, 2] = cx * cy
    return R


def rotation_matrix_to_rpy(R):
    """(..., 3, 3) -> (..., 3): inverse of ``rpy_to_rot_mat`` (R = Rx Ry Rz), the angles the
    reference's ``fk_rpy`` extracts (``src/acrobotics/robot.py:71-80``, including its
    ``r_z = atan2(-R01, R22)`` - exact whenever r_x = 0, which is the case the reference's
    tests cover, ``tests/test_path.py:139-176``).  ``acrolib.geometry.rotation_matrix_to_rpy``
    itself is un-vendored: for r_x != 0 the textbook ``atan2(-R01, R00)`` is used."""
    R = np.asarray(R, dtype=np.float64)
    r_x = np.arctan2(-R[..., 1, 2], R[..., 2, 2])
    r_y = np.arctan2(R[..., 0, 2], np.sqrt(R[..., 0, 0] ** 2 + R[..., 0, 1] ** 2))
    r_z = np.arctan2(-R[..., 0, 1], R[..., 0, 0])
    return np.stack([r_x, r_y, r_z], axis=-1)


def is_in_range(x, lower, upper):
    return not (x > upper or x < lower)


def check_rxyz_input(rxyz):
    """reference path/util.py:11-20"""
    if np.allclose(abs(rxyz[1]), np.pi / 2):
        return False
    return (is_in_range(rxyz[0], -np.pi, np.pi) and is_in_range(rxyz[1], -np.pi / 2, np.pi / 2)
            and is_in_range(rxyz[2], -np.pi, np.pi))
