"""Path points with position / Euler-angle tolerances and their joint solutions.

Reference: ``TolPositionPt`` (``src/acrobotics/path/path_pt.py:27-87``), ``TolEulerPt``
(``:90-147``), ``PathPt.to_joint_solutions`` with the GRID strategy
(``path/path_pt_base.py:87-111``) and ``path_to_joint_solutions``
(``planning/sampling_based.py:94-106``).  The reference holds the reference orientation as a
pyquaternion ``Quaternion``; here it is a 3x3 rotation matrix (or anything with a
``rotation_matrix`` attribute).

``joint_solutions_csr`` evaluates a whole path in one pipeline on the GPU
(``acro_joint_solutions_csr_f64``): every sampled pose of every path point goes through the
analytical IK, the filtered FK -> collision kernel and a CSR compaction.
"""
import ctypes

import numpy as np
import torch

from .. import _native, device
from ..transforms import tf_inverse
from .ranges import QuaternionTolerance, Tolerance
from .sampling import SampleMethod, Sampler, SamplingSetting, SearchStrategy
from .util import create_grid, rotation_matrix_to_rpy, rpy_to_rot_mat


def tf_inverse_batch(T):
    """Inverse of one rigid transform, shaped to broadcast against (..., 4, 4) batches."""
    return tf_inverse(T)


def _rotation(quat_or_matrix):
    R = getattr(quat_or_matrix, "rotation_matrix", quat_or_matrix)
    R = np.asarray(R, dtype=np.float64)
    if R.shape != (3, 3):
        raise ValueError("orientation must be a 3x3 rotation matrix or expose .rotation_matrix")
    return R


def _is_kuka(robot):
    """The fused IK -> collision -> CSR pipeline is the Kuka's closed form; every other robot
    goes through its own ``ik`` (reference path/path_pt_base.py:76-85)."""
    from ..catalog import Kuka

    return isinstance(robot, Kuka)


class _PathPt:
    """``Pose`` + ``PathPt`` of the reference (path/path_pt_base.py:15-129)."""

    tol = None
    sampler = None

    @property
    def rotation_matrix(self):
        return self._R

    @property
    def quat(self):
        """The reference keeps the orientation as a quaternion object; here only its
        ``rotation_matrix`` is ever needed, so the point itself plays that role."""
        return self

    @property
    def transformation_matrix(self):
        tf = np.eye(4)
        tf[:3, :3] = self._R
        tf[:3, 3] = self.pos
        return tf

    def translate(self, trans_vec):
        self.pos = self.pos + np.asarray(trans_vec, dtype=np.float64)

    def rotate(self, R):
        self._R = np.asarray(R, dtype=np.float64) @ self._R

    @staticmethod
    def count_tolerance(values):
        return sum(bool(v.has_tolerance) for v in values)

    # ---- sampling -----------------------------------------------------------------
    def sample_grid_array(self) -> np.ndarray:
        """All grid samples as one (S, 4, 4) array."""
        raise NotImplementedError

    def sample_grid(self):
        return list(self.sample_grid_array())

    def _relative_incremental(self, num_samples, method, tols):
        """(num_samples, len(tols)) values inside the tolerance ranges: unit-cube samples
        scaled to ``[lower, upper]``, zeros where there is no tolerance
        (path/path_pt.py:69-85)."""
        R = self.sampler.sample(num_samples, self.sample_dim, method)
        rel = np.zeros((num_samples, len(tols)))
        cnt = 0
        for i, value in enumerate(tols):
            if value.has_tolerance:
                rel[:, i] = R[:, cnt] * (value.upper - value.lower) + value.lower
                cnt += 1
        return rel

    def sample_incremental_array(self, num_samples, method) -> np.ndarray:
        raise NotImplementedError

    def sample_incremental(self, num_samples, method):
        return list(self.sample_incremental_array(num_samples, method))

    def transform_to_rel_tolerance_deviation(self, tf):
        raise NotImplementedError

    # ---- joint solutions ------------------------------------------------------------
    @staticmethod
    def _calc_ik(robot, samples):
        """All IK solutions of a list of poses, one ``robot.ik`` call per pose
        (path/path_pt_base.py:76-85) - the generic route for robots without the fused
        pipeline."""
        joint_solutions = []
        for transform in samples:
            ik_result = robot.ik(transform)
            if ik_result.success:
                joint_solutions.extend(ik_result.solutions)
        return joint_solutions

    def _free_solutions(self, robot, samples, scene) -> np.ndarray:
        """Collision-free IK solutions of the (S, 4, 4) poses ``samples`` in the reference's
        order.  Kuka: one fused launch; other robots: their own ``ik`` per pose, then one
        batched collision call."""
        samples = np.asarray(samples, dtype=np.float64).reshape(-1, 4, 4)
        if _is_kuka(robot):
            q, offsets = joint_solutions_csr(robot, samples[None], scene)
            return q[offsets[0]:offsets[1]]
        js = self._calc_ik(robot, list(samples))
        if not js:
            return np.zeros((0, robot.ndof))
        js = np.array(js, dtype=np.float64).reshape(len(js), -1)
        if hasattr(robot, "is_in_collision_batch"):
            if scene is not None:
                hit = np.asarray(robot.is_in_collision_batch(js, scene))
            else:
                hit = np.asarray(robot.is_in_self_collision_batch(js))
        else:  # duck-typed robots (the reference's tests use a stub with is_in_collision only)
            hit = np.array([bool(robot.is_in_collision(q, scene)) for q in js])
        return js[~hit]

    def to_joint_solutions(self, robot, settings=None, scene=None) -> np.ndarray:
        """All collision-free joint positions that satisfy the constraints
        (path/path_pt_base.py:87-111): (M, ndof) array, M may be 0.  ``settings=None`` means
        the GRID strategy."""
        strategy = SearchStrategy.GRID if settings is None else settings.search_strategy
        if strategy == SearchStrategy.GRID:
            return self._free_solutions(robot, self.sample_grid_array(), scene)
        if strategy == SearchStrategy.INCREMENTAL:
            samples = self.sample_incremental_array(settings.num_samples, settings.sample_method)
            return self._free_solutions(robot, samples, scene)
        if strategy == SearchStrategy.MIN_INCREMENTAL:
            return self._incremental_search(robot, scene, settings)
        raise NotImplementedError(strategy)

    def _incremental_search(self, robot, scene, s):
        """Sample ``step_size`` poses at a time until ``desired_num_samples`` collision-free
        joint solutions are found (path/path_pt_base.py:113-129)."""
        found = []
        total = 0
        for _ in range(s.max_search_iters):
            samples = self.sample_incremental_array(s.step_size, s.sample_method)
            js = self._free_solutions(robot, samples, scene)
            if len(js):
                found.append(js)
                total += len(js)
            if total >= s.desired_num_samples:
                return np.concatenate(found)
        raise Exception("Maximum iterations reached in to_joint_solutions.")


class TolPositionPt(_PathPt):
    def __init__(self, position, quaternion, tolerance):
        self.pos = np.array(position, dtype=np.float64)
        self._R = _rotation(quaternion)
        self.pos_tol = tolerance
        self.tol = tolerance
        self.sampler = Sampler()
        self.sample_dim = self.count_tolerance(tolerance)

    def to_transform(self, sampled_pos):
        tf = np.eye(4)
        tf[:3, :3] = self._R
        tf[:3, 3] = np.array(sampled_pos)
        return tf

    def transform_to_rel_tolerance_deviation(self, tf):
        """Pose in the world frame -> position in the local path frame (path_pt.py:42-45)."""
        T_rel = tf_inverse(self.transformation_matrix) @ np.asarray(tf, dtype=np.float64)
        return T_rel[:3, 3]

    def rel_to_abs(self, grid):
        grid = np.asarray(grid, dtype=np.float64)
        return self.pos + grid[:, :3] @ self._R.T

    def sample_relative_grid(self):
        return create_grid([tol.discretize() for tol in self.tol])

    def _poses(self, pos):
        tf = np.tile(np.eye(4), (len(pos), 1, 1))
        tf[:, :3, :3] = self._R
        tf[:, :3, 3] = pos
        return tf

    def sample_grid_array(self):
        return self._poses(self.rel_to_abs(self.sample_relative_grid()))

    def sample_incremental_array(self, num_samples, method):
        return self._poses(self.rel_to_abs(self._relative_incremental(num_samples, method, self.tol)))


class TolEulerPt(_PathPt):
    def __init__(self, position, quaternion, pos_tol, rot_tol):
        self.pos = np.array(position, dtype=np.float64)
        self._R = _rotation(quaternion)
        self.pos_tol = pos_tol
        self.rot_tol = rot_tol
        self.tol = pos_tol + rot_tol
        self.sample_dim = self.count_tolerance(self.tol)
        self.sampler = Sampler()

    def to_transform(self, sample):
        sample = np.asarray(sample, dtype=np.float64)
        tf = np.eye(4)
        tf[:3, 3] = sample[:3]
        tf[:3, :3] = self._R @ rpy_to_rot_mat(sample[3:])
        return tf

    def transform_to_rel_tolerance_deviation(self, tf):
        """Pose in the world frame -> (x, y, z, r_x, r_y, r_z) in the local path frame
        (path_pt.py:110-115); also accepts an (N, 4, 4) batch."""
        T_rel = tf_inverse_batch(self.transformation_matrix) @ np.asarray(tf, dtype=np.float64)
        return np.concatenate([T_rel[..., :3, 3], rotation_matrix_to_rpy(T_rel[..., :3, :3])], axis=-1)

    def rel_to_abs(self, grid):
        grid = np.asarray(grid, dtype=np.float64)
        tf = np.tile(np.eye(4), (len(grid), 1, 1))
        tf[:, :3, 3] = self.pos + grid[:, :3] @ self._R.T
        tf[:, :3, :3] = self._R @ rpy_to_rot_mat(grid[:, 3:6])
        return tf

    def sample_grid_array(self):
        return self.rel_to_abs(create_grid([tol.discretize() for tol in self.tol]))

    def sample_incremental_array(self, num_samples, method):
        return self.rel_to_abs(self._relative_incremental(num_samples, method, self.tol))


def _random_rotation_near(rng, n, max_quat_distance):
    """n rotations whose quaternion distance to the identity is at most ``max_quat_distance``
    (pyquaternion's ``Quaternion.distance`` = |log(q0^-1 q1)| = half the rotation angle): a
    uniformly random axis and an angle uniform in ``[0, 2 d]``.  The reference calls
    ``acrolib.quaternion.Quaternion.random_near`` (un-vendored; its distribution is unpinned)
    and its test only asserts the distance bound (``tests/test_path.py:189-199``)."""
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    angle = rng.uniform(0.0, 2.0 * min(float(max_quat_distance), np.pi / 2), n)
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    s, c = np.sin(angle)[:, None, None], np.cos(angle)[:, None, None]
    return np.eye(3) + s * K + (1.0 - c) * (K @ K)


class TolQuatPt(_PathPt):
    """Position tolerances plus a ball of orientations (path_pt.py:150-225)."""

    def __init__(self, position, quaternion, pos_tol, quat_tol):
        self.pos = np.array(position, dtype=np.float64)
        self._R = _rotation(quaternion)
        self.pos_tol = pos_tol
        self.quat_tol = quat_tol
        self.tol = pos_tol + [quat_tol]
        self.sample_dim = self.count_tolerance(self.pos_tol)
        self.sampler = Sampler()

    def to_transform(self, pos_sample, quat_sample):
        tf = np.eye(4)
        tf[:3, 3] = pos_sample
        tf[:3, :3] = _rotation(quat_sample)
        return tf

    def transform_to_rel_tolerance_deviation(self, tf):
        tf = np.asarray(tf, dtype=np.float64)
        T_rel = tf_inverse(self.transformation_matrix) @ tf
        c = 0.5 * (np.trace(T_rel[:3, :3]) - 1.0)
        return np.hstack((T_rel[:3, 3], 0.5 * np.arccos(np.clip(c, -1.0, 1.0))))

    def rel_to_abs(self, pos_grid):
        return self.pos + np.asarray(pos_grid, dtype=np.float64)[:, :3] @ self._R.T

    def sample_grid_array(self):
        raise NotImplementedError  # as in the reference (path_pt.py:196-197)

    def sample_incremental_array(self, num_samples, method):
        pos = self.rel_to_abs(self._relative_incremental(num_samples, method, self.pos_tol))
        if method != SampleMethod.random_uniform:
            raise NotImplementedError
        Rn = _random_rotation_near(self.sampler.rng, num_samples, self.quat_tol.dist)
        tf = np.tile(np.eye(4), (num_samples, 1, 1))
        tf[:, :3, :3] = self._R @ Rn
        tf[:, :3, 3] = pos
        return tf


class FreeOrientationPt(TolQuatPt):
    def __init__(self, position, pos_tol):
        super().__init__(position, np.eye(3), pos_tol, QuaternionTolerance(0.25 * np.pi + 0.01))


def _tolerance_key(pt):
    return (type(pt), tuple((t.lower, t.upper, t.num_samples, t.has_tolerance) for t in pt.tol))


def path_grid_arrays(path):
    """Grid samples of a whole path without a Python loop per point: points of the same kind
    with the same tolerances share their relative grid, so their (S, 4, 4) sample arrays are
    one broadcast.  Returns a list of (indices, T) with T of shape (len(indices), S, 4, 4)."""
    groups = {}
    for i, pt in enumerate(path):
        groups.setdefault(_tolerance_key(pt), []).append(i)
    out = []
    for idxs in groups.values():
        first = path[idxs[0]]
        if not isinstance(first, (TolPositionPt, TolEulerPt)):
            out.append((idxs, np.stack([path[i].sample_grid_array() for i in idxs])))
            continue
        grid = create_grid([tol.discretize() for tol in first.tol]).astype(np.float64)  # (S, 3|6)
        R = np.stack([path[i].rotation_matrix for i in idxs])                          # (P, 3, 3)
        pos = np.stack([path[i].pos for i in idxs])                                    # (P, 3)
        P, S = len(idxs), len(grid)
        T = np.zeros((P, S, 4, 4))
        T[..., 3, 3] = 1.0
        T[..., :3, 3] = pos[:, None, :] + np.einsum("sj,pij->psi", grid[:, :3], R)
        if isinstance(first, TolEulerPt):
            T[..., :3, :3] = np.einsum("pij,sjk->psik", R, rpy_to_rot_mat(grid[:, 3:6]))
        else:
            T[..., :3, :3] = R[:, None]
        out.append((idxs, T))
    return out


def joint_solutions_csr(robot, T, scene, return_all=False, trim=True):
    """Planner inner loop for a whole path.

    ``T``: (P, S, 4, 4) sampled poses (P path points, S samples each; numpy or CUDA tensor).
    Returns ``(q_free, offsets)``: ``q_free[offsets[p]:offsets[p+1]]`` are the collision-free
    IK solutions of path point ``p`` in the reference's order.  numpy in -> numpy out, CUDA
    tensor in -> CUDA tensors out.  ``return_all`` also returns (q_all (P,S,8,6), valid (P,S),
    free (P,S)) branch bit masks.  ``trim=False`` (CUDA input only) skips the host
    synchronisation that reads the row count: ``q_free`` then has the full capacity of ``8 P S``
    rows, of which the first ``offsets[-1]`` are defined - the call stays asynchronous."""
    host = not isinstance(T, torch.Tensor) or not T.is_cuda
    device.require_cuda()
    Td = torch.as_tensor(np.asarray(T) if not isinstance(T, torch.Tensor) else T,
                         dtype=torch.float64).to("cuda").contiguous()
    if Td.dim() != 4 or Td.shape[-2:] != (4, 4):
        raise ValueError("T must have shape (P, S, 4, 4)")
    P, S = int(Td.shape[0]), int(Td.shape[1])
    n = P * S
    q_all = torch.empty((max(n, 1), 8, 6), dtype=torch.float64, device="cuda")
    valid = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    free = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    q_free = torch.empty((max(8 * n, 1), 6), dtype=torch.float64, device="cuda")
    offsets = torch.zeros(P + 1, dtype=torch.int64, device="cuda")
    robot.cc_checks += 8 * n
    _native.check(
        _native.load().acro_joint_solutions_csr_f64(
            robot._handle(), scene.native_handle() if scene is not None else None,
            Td.data_ptr(), P, max(S, 1), q_all.data_ptr(), valid.data_ptr(), free.data_ptr(),
            q_free.data_ptr(), 8 * n, offsets.data_ptr(), device.current_stream_ptr(),
        ) if S > 0 else 0,
        "acro_joint_solutions_csr_f64",
    )
    if trim or host:
        total = int(offsets[-1].item())
        q_free = q_free[:total]
    out = (device.to_host(q_free, host), device.to_host(offsets, host))
    if return_all:
        out += (device.to_host(q_all[:n].reshape(P, S, 8, 6), host),
                device.to_host(valid[:n].reshape(P, S), host),
                device.to_host(free[:n].reshape(P, S), host))
    return out


def path_to_joint_solutions(path, robot, scene=None, settings=None):
    """List of (M_i, ndof) arrays, one per path point (planning/sampling_based.py:94-106);
    raises like the reference when a point has no valid joint solution.

    Kuka + GRID / INCREMENTAL: the whole path goes through the fused GPU pipeline - points of
    the same kind with the same tolerances form one launch (GRID), respectively every point
    draws its ``num_samples`` poses and all of them form one launch (INCREMENTAL).
    MIN_INCREMENTAL: all points search side by side, one launch per search iteration, and a
    point stops sampling once it has ``desired_num_samples`` solutions (the per-point
    semantics of path/path_pt_base.py:113-129).  Other robots: point by point through
    ``PathPt.to_joint_solutions``."""
    strategy = SearchStrategy.GRID if settings is None else settings.search_strategy
    out = [None] * len(path)
    if not _is_kuka(robot):
        out = [pt.to_joint_solutions(robot, settings, scene) for pt in path]
    elif strategy == SearchStrategy.GRID:
        for idxs, T in path_grid_arrays(path):
            q, off = joint_solutions_csr(robot, T, scene)
            for k, i in enumerate(idxs):
                out[i] = q[off[k]:off[k + 1]]
    elif strategy == SearchStrategy.INCREMENTAL:
        T = np.stack([pt.sample_incremental_array(settings.num_samples, settings.sample_method)
                      for pt in path])
        q, off = joint_solutions_csr(robot, T, scene)
        out = [q[off[k]:off[k + 1]] for k in range(len(path))]
    elif strategy == SearchStrategy.MIN_INCREMENTAL:
        found = [[] for _ in path]
        count = np.zeros(len(path), dtype=np.int64)
        active = list(range(len(path)))
        for _ in range(settings.max_search_iters):
            T = np.stack([path[i].sample_incremental_array(settings.step_size, settings.sample_method)
                          for i in active])
            q, off = joint_solutions_csr(robot, T, scene)
            for k, i in enumerate(active):
                if off[k + 1] > off[k]:
                    found[i].append(q[off[k]:off[k + 1]])
                    count[i] += off[k + 1] - off[k]
            active = [i for i in active if count[i] < settings.desired_num_samples]
            if not active:
                break
        if active:
            raise Exception("Maximum iterations reached in to_joint_solutions.")
        out = [np.concatenate(f) for f in found]
    else:
        raise NotImplementedError(strategy)
    for i, js in enumerate(out):
        if len(js) == 0:
            raise Exception(f"No valid joint solutions for path point {i}.")
    return out
