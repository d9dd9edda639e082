"""Copies of a toleranced path point along a line, a circle or an arc (reference
``src/acrobotics/path/factory.py:11-58``).  The reference turns the arc angle into a rotation
matrix with pyquaternion; here it is Rodrigues' formula."""
from copy import deepcopy

import numpy as np


def check_num_points(num_points: int):
    if num_points < 2:
        raise Exception(f"Value of num_points must be 2 or more, not {num_points}.")


def axis_angle_matrix(axis, angle):
    k = np.asarray(axis, dtype=np.float64)
    k = k / np.linalg.norm(k)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


def create_line(start_pt, end_pos, num_points: int):
    """Copy a toleranced path point along a straight line (the first element is
    ``start_pt`` itself, as in the reference)."""
    check_num_points(num_points)
    trans_vec = (np.asarray(end_pos, dtype=np.float64) - start_pt.pos) / (num_points - 1)
    path = [start_pt]
    for _ in range(num_points - 1):
        new_pt = deepcopy(path[-1])
        new_pt.translate(trans_vec)
        path.append(new_pt)
    return path


def create_arc(start_pt, mid_point, rotation_axis, angle, num_points):
    check_num_points(num_points)
    rotating_vector = start_pt.pos - np.asarray(mid_point, dtype=np.float64)
    path = [deepcopy(start_pt)]
    for ai in np.linspace(angle / (num_points - 1), angle, num_points - 1):
        rot_mat = axis_angle_matrix(rotation_axis, ai)
        new_pt = deepcopy(start_pt)
        new_pt.translate(rot_mat @ rotating_vector - rotating_vector)
        new_pt.rotate(rot_mat)
        path.append(new_pt)
    return path


def create_circle(start_pt, mid_point, rotation_axis, num_points: int):
    check_num_points(num_points)
    return create_arc(start_pt, mid_point, rotation_axis, 2 * np.pi, num_points)
