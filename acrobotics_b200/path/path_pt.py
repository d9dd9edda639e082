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
from .ranges import Tolerance
from .util import create_grid, rpy_to_rot_mat


def _rotation(quat_or_matrix):
    R = getattr(quat_or_matrix, "rotation_matrix", quat_or_matrix)
    R = np.asarray(R, dtype=np.float64)
    if R.shape != (3, 3):
        raise ValueError("orientation must be a 3x3 rotation matrix or expose .rotation_matrix")
    return R


class _PathPt:
    tol = None

    @property
    def rotation_matrix(self):
        return self._R

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

    def sample_grid_array(self) -> np.ndarray:
        """All grid samples as one (S, 4, 4) array."""
        raise NotImplementedError

    def sample_grid(self):
        return list(self.sample_grid_array())

    def to_joint_solutions(self, robot, settings=None, scene=None) -> np.ndarray:
        """All collision-free joint positions that satisfy the constraints (GRID strategy):
        (M, ndof) array, M may be 0."""
        q, offsets = joint_solutions_csr(robot, self.sample_grid_array()[None], scene)
        return q[offsets[0]:offsets[1]]


class TolPositionPt(_PathPt):
    def __init__(self, position, quaternion, tolerance):
        self.pos = np.array(position, dtype=np.float64)
        self._R = _rotation(quaternion)
        self.pos_tol = tolerance
        self.tol = tolerance
        self.sample_dim = self.count_tolerance(tolerance)

    def rel_to_abs(self, grid):
        grid = np.asarray(grid, dtype=np.float64)
        return self.pos + grid[:, :3] @ self._R.T

    def sample_relative_grid(self):
        return create_grid([tol.discretize() for tol in self.tol])

    def sample_grid_array(self):
        pos = self.rel_to_abs(self.sample_relative_grid())
        tf = np.tile(np.eye(4), (len(pos), 1, 1))
        tf[:, :3, :3] = self._R
        tf[:, :3, 3] = pos
        return tf


class TolEulerPt(_PathPt):
    def __init__(self, position, quaternion, pos_tol, rot_tol):
        self.pos = np.array(position, dtype=np.float64)
        self._R = _rotation(quaternion)
        self.pos_tol = pos_tol
        self.rot_tol = rot_tol
        self.tol = pos_tol + rot_tol
        self.sample_dim = self.count_tolerance(self.tol)

    def rel_to_abs(self, grid):
        grid = np.asarray(grid, dtype=np.float64)
        tf = np.tile(np.eye(4), (len(grid), 1, 1))
        tf[:, :3, 3] = self.pos + grid[:, :3] @ self._R.T
        tf[:, :3, :3] = self._R @ rpy_to_rot_mat(grid[:, 3:6])
        return tf

    def sample_grid_array(self):
        return self.rel_to_abs(create_grid([tol.discretize() for tol in self.tol]))


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


def joint_solutions_csr(robot, T, scene, return_all=False):
    """Planner inner loop for a whole path.

    ``T``: (P, S, 4, 4) sampled poses (P path points, S samples each; numpy or CUDA tensor).
    Returns ``(q_free, offsets)``: ``q_free[offsets[p]:offsets[p+1]]`` are the collision-free
    IK solutions of path point ``p`` in the reference's order.  numpy in -> numpy out, CUDA
    tensor in -> CUDA tensors out.  ``return_all`` also returns (q_all (P,S,8,6), valid (P,S),
    free (P,S)) branch bit masks."""
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
    total = int(offsets[-1].item())
    q_free = q_free[:total]
    out = (device.to_host(q_free, host), device.to_host(offsets, host))
    if return_all:
        out += (device.to_host(q_all[:n].reshape(P, S, 8, 6), host),
                device.to_host(valid[:n].reshape(P, S), host),
                device.to_host(free[:n].reshape(P, S), host))
    return out


def path_to_joint_solutions(path, robot, scene=None):
    """List of (M_i, ndof) arrays, one per path point (planning/sampling_based.py:94-106);
    raises like the reference when a point has no valid joint solution.  Points may have
    different tolerances; points with the same tolerances form one launch."""
    out = [None] * len(path)
    for idxs, T in path_grid_arrays(path):
        q, off = joint_solutions_csr(robot, T, scene)
        for k, i in enumerate(idxs):
            out[i] = q[off[k]:off[k + 1]]
    for i, js in enumerate(out):
        if len(js) == 0:
            raise Exception(f"No valid joint solutions for path point {i}.")
    return out
