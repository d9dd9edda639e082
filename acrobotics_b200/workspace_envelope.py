"""Workspace envelope: fraction of reachable, self-collision-free IK solutions over a grid
of positions (reference ``src/acrobotics/workspace_envelope.py:9-85``).

The reference loops over grid points x random orientations and calls ``robot.ik`` and
``robot.is_in_self_collision`` per solution; here all poses of the grid go through the batched
IK -> self-collision filter pipeline in one call.  The random orientations (pyquaternion's
``Quaternion.random()`` in the reference) come from a seedable numpy generator: the sampling
is statistically, not bit-wise, the same."""
import numpy as np


class EnvelopeSettings:
    def __init__(self, sample_distance: float, num_orientation_samples: int, max_ik_solutions: int):
        self.sample_distance = sample_distance
        self.num_orientation_samples = num_orientation_samples
        self.max_ik_solutions = max_ik_solutions


def random_rotations(n, rng):
    """n rotation matrices, uniform on SO(3) (normalised Gaussian quaternions)."""
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    w, x, y, z = quat.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def sample_position(position, n, rng=None):
    """n random transforms at the given position (workspace_envelope.py:21-26)."""
    rng = np.random.default_rng() if rng is None else rng
    tf = np.tile(np.eye(4), (n, 1, 1))
    tf[:, :3, :3] = random_rotations(n, rng)
    tf[:, :3, 3] = position
    return list(tf)


def generate_positions(max_extension, num_points: int):
    """workspace_envelope.py:66-70."""
    steps = np.linspace(-max_extension, max_extension, num_points)
    x, y, z = np.meshgrid(steps, steps, steps)
    return np.vstack((x.flatten(), y.flatten(), z.flatten())).T


def calculate_reachability_batch(robot, positions, num_samples=100, max_ik_solutions=8, rng=None):
    """Reachability of every position (P, 3): collision-free IK solutions over
    ``max_ik_solutions * num_samples`` (workspace_envelope.py:42-57), one GPU pipeline."""
    rng = np.random.default_rng(0) if rng is None else rng
    positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
    P = len(positions)
    T = np.tile(np.eye(4), (P, num_samples, 1, 1))
    T[:, :, :3, :3] = random_rotations(P * num_samples, rng).reshape(P, num_samples, 3, 3)
    T[:, :, :3, 3] = positions[:, None, :]
    _, _, free = robot.ik_filtered_batch(T.reshape(-1, 4, 4), None, _force_self=True)
    counts = free.reshape(P, num_samples * 8).sum(axis=1)
    return counts / (max_ik_solutions * num_samples)


def calculate_reachability(robot, position, num_samples=100, max_ik_solutions=8, rng=None):
    return float(calculate_reachability_batch(robot, np.asarray(position)[None], num_samples,
                                              max_ik_solutions, rng)[0])


def generate_robot_envelope(robot, settings: EnvelopeSettings, rng=None):
    """(num_points^3, 4) rows [reachability, x, y, z] (workspace_envelope.py:73-85)."""
    max_extension = robot.estimate_max_extension()
    num_points = int(2 * max_extension / settings.sample_distance)
    points = generate_positions(max_extension, num_points)
    result = np.zeros((len(points), 4))
    result[:, 0] = calculate_reachability_batch(
        robot, points, settings.num_orientation_samples, settings.max_ik_solutions, rng)
    result[:, 1:] = points
    return result
