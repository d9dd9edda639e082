"""Homogeneous-transform helpers the reference takes from ``acrolib.geometry``
(``translation``, ``pose_x`` ... used by reference
``src/acrobotics/robot_examples.py:2`` and ``tool_examples.py:5`` to build the
constant shape offsets) plus ``tf_inverse`` (``robot_examples.py:14-23``).
Host-side constants only - nothing here is on the per-configuration path.
"""
import numpy as np


def _rot(axis, angle):
    c, s = np.cos(angle), np.sin(angle)
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    R = np.eye(3)
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def rot_x(angle):
    return _rot(0, angle)


def rot_y(angle):
    return _rot(1, angle)


def rot_z(angle):
    return _rot(2, angle)


def _pose(axis, angle, x, y, z):
    tf = np.eye(4)
    tf[:3, :3] = _rot(axis, angle)
    tf[:3, 3] = (x, y, z)
    return tf


def pose_x(angle, x, y, z):
    return _pose(0, angle, x, y, z)


def pose_y(angle, x, y, z):
    return _pose(1, angle, x, y, z)


def pose_z(angle, x, y, z):
    return _pose(2, angle, x, y, z)


def translation(x, y, z):
    return _pose(0, 0.0, x, y, z)


def tf_inverse(T):
    """Inverse of a rigid transform (copy)."""
    T = np.asarray(T, dtype=np.float64)
    Ti = np.eye(4)
    Ti[:3, :3] = T[:3, :3].T
    Ti[:3, 3] = -Ti[:3, :3] @ T[:3, 3]
    return Ti


def pack_3x4(tf):
    """4x4 (or 3x4) array -> 12 floats, row-major 3x4: the ABI's transform format."""
    return np.asarray(tf, dtype=np.float64)[:3, :4].reshape(12)
