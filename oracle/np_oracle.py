"""Vectorised FP64 numpy restatement of the acrobotics hot path (TEST INFRASTRUCTURE).

Every function cites the reference code it follows (paths relative to
``/root/reference``).  All arrays are float64; ``q`` is ``(N, ndof)``.

The robot is described by a ``RobotSpec`` (plain arrays), obtained with
``spec_from_robot`` from either a *reference* robot object (through
``oracle.ref_shim``) or a product ``acrobotics_b200`` robot object - both expose
the same attributes (``links[i].dh``, ``links[i].joint_type``,
``links[i].geometry.s/.tf_s``, ``tf_base``, ``tf_tool_tip``, ``geometry_tool``,
``geometry_base``, ``collision_matrix``, ``do_check_self_collision``).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

FUDGE2 = 1.0e-6
EPS = float(np.finfo(np.float64).eps)

GROUP_BASE = -1  # box rides on tf_base            (robot.py:257-260)
GROUP_TOOL = -2  # box rides on the last link frame (robot.py:251-254, :214-220)


@dataclass
class BoxSpec:
    group: int  # link index, GROUP_BASE or GROUP_TOOL
    half: np.ndarray  # (3,) half extents
    tf: np.ndarray  # (4,4) offset of the box in its carrier frame


@dataclass
class RobotSpec:
    dh: np.ndarray  # (ndof, 4): a, alpha, d, theta   (link.py:8)
    prismatic: np.ndarray  # (ndof,) bool                     (link.py:11-13)
    tf_base: np.ndarray  # (4,4)                               (robot.py:55)
    tf_tool_tip: Optional[np.ndarray]  # (4,4) or None          (robot.py:57)
    boxes: List[BoxSpec] = field(default_factory=list)
    self_pairs: List[Tuple[int, int]] = field(default_factory=list)  # groups (gi, gj)
    check_self: bool = True

    @property
    def ndof(self):
        return self.dh.shape[0]


def spec_from_robot(robot) -> RobotSpec:
    """Flatten a reference-style robot object (duck typed)."""
    ndof = robot.ndof
    dh = np.array(
        [[l.dh.a, l.dh.alpha, l.dh.d, l.dh.theta] for l in robot.links], dtype=np.float64
    )
    prismatic = np.array(
        [getattr(l.joint_type, "value", l.joint_type) == "p" for l in robot.links], dtype=bool
    )
    boxes = []
    for i, l in enumerate(robot.links):
        for s, tf in zip(l.geometry.s, l.geometry.tf_s):
            boxes.append(BoxSpec(i, 0.5 * np.array([s.dx, s.dy, s.dz], float), np.array(tf, float)))
    tool = getattr(robot, "geometry_tool", None)
    if tool is not None:
        for s, tf in zip(tool.s, tool.tf_s):
            boxes.append(
                BoxSpec(GROUP_TOOL, 0.5 * np.array([s.dx, s.dy, s.dz], float), np.array(tf, float))
            )
    base = getattr(robot, "geometry_base", None)
    if base is not None:
        for s, tf in zip(base.s, base.tf_s):
            boxes.append(
                BoxSpec(GROUP_BASE, 0.5 * np.array([s.dx, s.dy, s.dz], float), np.array(tf, float))
            )
    # robot.py:206-211 visits every ordered (i, j) with collision_matrix[i, j];
    # the predicate is symmetric, so unique unordered pairs suffice.
    cm = np.asarray(robot.collision_matrix, dtype=bool)
    pairs = [(i, j) for i in range(ndof) for j in range(i + 1, ndof) if cm[i, j] or cm[j, i]]
    # robot.py:214-220: tool against every link but the last one
    if tool is not None:
        pairs += [(i, GROUP_TOOL) for i in range(ndof - 1)]
    tip = getattr(robot, "tf_tool_tip", None)
    return RobotSpec(
        dh=dh,
        prismatic=prismatic,
        tf_base=np.array(robot.tf_base, float),
        tf_tool_tip=None if tip is None else np.array(tip, float),
        boxes=boxes,
        self_pairs=pairs,
        check_self=bool(robot.do_check_self_collision),
    )


def scene_arrays(scene):
    """``Scene`` (geometry.py:12-14) -> (half (S,3), R (S,3,3), T (S,3))."""
    half = np.array([[0.5 * s.dx, 0.5 * s.dy, 0.5 * s.dz] for s in scene.s], dtype=np.float64)
    tf = np.array([np.asarray(t, dtype=np.float64) for t in scene.tf_s]).reshape(-1, 4, 4)
    return half.reshape(-1, 3), tf[:, :3, :3].copy(), tf[:, :3, 3].copy()


# ----------------------------------------------------------------------------
# forward kinematics
# ----------------------------------------------------------------------------
def dh_transforms(spec: RobotSpec, q) -> np.ndarray:
    """(N, ndof, 4, 4) link-relative DH transforms; link.py:35-58."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    N, ndof = q.shape
    assert ndof == spec.ndof
    a = np.broadcast_to(spec.dh[:, 0], (N, ndof))
    alpha = np.broadcast_to(spec.dh[:, 1], (N, ndof))
    d = np.where(spec.prismatic, q, spec.dh[:, 2])  # prismatic: d = q   (link.py:43-44)
    theta = np.where(spec.prismatic, spec.dh[:, 3], q)  # revolute: theta = q (link.py:41-42)
    ct, st = np.cos(theta), np.sin(theta)
    ca, sa = np.cos(alpha), np.sin(alpha)  # evaluated per call in the reference (link.py:48-49)
    T = np.zeros((N, ndof, 4, 4))
    T[..., 0, 0], T[..., 0, 1], T[..., 0, 2], T[..., 0, 3] = ct, -st * ca, st * sa, a * ct
    T[..., 1, 0], T[..., 1, 1], T[..., 1, 2], T[..., 1, 3] = st, ct * ca, -ct * sa, a * st
    T[..., 2, 1], T[..., 2, 2], T[..., 2, 3] = sa, ca, d
    T[..., 3, 3] = 1.0
    return T


def fk_all_links(spec: RobotSpec, q) -> np.ndarray:
    """(N, ndof, 4, 4) world frames of every link, base included, tool tip
    excluded; robot.py:82-91."""
    Ti = dh_transforms(spec, q)
    N, ndof = Ti.shape[:2]
    out = np.empty_like(Ti)
    T = np.broadcast_to(spec.tf_base, (N, 4, 4))
    for i in range(ndof):
        T = T @ Ti[:, i]
        out[:, i] = T
    return out


def fk(spec: RobotSpec, q) -> np.ndarray:
    """(N, 4, 4) end-effector pose incl. tool tip; robot.py:59-69."""
    T = fk_all_links(spec, q)[:, -1]
    if spec.tf_tool_tip is not None:
        T = T @ spec.tf_tool_tip
    return T


def rpy_of(T) -> np.ndarray:
    """The reference's 6-vector [x,y,z,r_x,r_y,r_z]; robot.py:71-80 (note the
    authoritative quirk r_z = atan2(-T01, T22))."""
    s = np.sqrt(T[..., 0, 0] * T[..., 0, 0] + T[..., 0, 1] * T[..., 0, 1])
    out = np.empty(T.shape[:-2] + (6,))
    out[..., :3] = T[..., :3, 3]
    out[..., 3] = np.arctan2(-T[..., 1, 2], T[..., 2, 2])
    out[..., 4] = np.arctan2(T[..., 0, 2], s)
    out[..., 5] = np.arctan2(-T[..., 0, 1], T[..., 2, 2])
    return out


def fk_rpy(spec: RobotSpec, q) -> np.ndarray:
    return rpy_of(fk(spec, q))


# ----------------------------------------------------------------------------
# FCL box-box separating-axis test, vectorised (see oracle/fcl_boxbox.py)
# ----------------------------------------------------------------------------
def sat(A, R1, T1, B, R2, T2):
    """Broadcasting box-box test.  ``A, B`` half extents (...,3); ``R*`` (...,3,3);
    ``T*`` (...,3).  Returns ``(collide, g)`` with ``g`` the largest decision
    quantity over the 15 axes (> 0 <=> FCL finds a separating axis)."""
    p = T2 - T1
    R1t = np.swapaxes(R1, -1, -2)
    pp = np.einsum("...ij,...j->...i", R1t, p)
    R = R1t @ R2
    Q = np.abs(R)
    g = np.full(np.broadcast_shapes(pp.shape[:-1], R.shape[:-2]), -np.inf)
    # face axes of box 1
    s2 = np.abs(pp) - (np.einsum("...ij,...j->...i", Q, B) + A)
    g = np.maximum(g, s2.max(axis=-1))
    # face axes of box 2
    tmp = np.einsum("...ij,...i->...j", R2, p)
    s2 = np.abs(tmp) - (np.einsum("...ij,...i->...j", Q, A) + B)
    g = np.maximum(g, s2.max(axis=-1))
    # edge axes
    Qf = Q + FUDGE2
    for i in range(3):
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            tmp = pp[..., i2] * R[..., i1, j] - pp[..., i1] * R[..., i2, j]
            s2 = np.abs(tmp) - (
                A[..., i1] * Qf[..., i2, j]
                + A[..., i2] * Qf[..., i1, j]
                + B[..., j1] * Qf[..., i, j2]
                + B[..., j2] * Qf[..., i, j1]
            )
            g = np.maximum(g, s2 - EPS)
    return ~(g > 0), g


# ----------------------------------------------------------------------------
# robot vs scene + self collision
# ----------------------------------------------------------------------------
def _carrier_frames(spec: RobotSpec, frames: np.ndarray, group: int) -> np.ndarray:
    if group == GROUP_BASE:
        return np.broadcast_to(spec.tf_base, frames[:, 0].shape)
    if group == GROUP_TOOL:
        return frames[:, -1]
    return frames[:, group]


def world_boxes(spec: RobotSpec, q):
    """World rotation (N,nb,3,3) and centre (N,nb,3) of every robot box:
    ``tf_self @ tf`` (geometry.py:56-57) with tf_self the carrier frame."""
    frames = fk_all_links(spec, q)
    N = frames.shape[0]
    nb = len(spec.boxes)
    Rw = np.empty((N, nb, 3, 3))
    cw = np.empty((N, nb, 3))
    for b, box in enumerate(spec.boxes):
        W = _carrier_frames(spec, frames, box.group) @ box.tf
        Rw[:, b] = W[:, :3, :3]
        cw[:, b] = W[:, :3, 3]
    return Rw, cw


def collision_margins(spec: RobotSpec, scene, q, chunk=20000):
    """Per configuration: ``g_min`` = the smallest, over every box pair the
    reference would test, of that pair's largest SAT decision quantity.
    ``g_min <= 0``  <=>  ``Robot.is_in_collision(q, scene)`` (robot.py:244-273:
    OR over tool x scene, base x scene, link x scene and, when
    ``do_check_self_collision``, the self pairs of robot.py:206-221)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    N = q.shape[0]
    gmin = np.full(N, np.inf)
    if scene is not None:
        sh, sR, sT = scene if isinstance(scene, tuple) else scene_arrays(scene)
    halfs = np.array([b.half for b in spec.boxes]).reshape(-1, 3)
    groups = [b.group for b in spec.boxes]
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        Rw, cw = world_boxes(spec, q[lo:hi])
        g = np.full(hi - lo, np.inf)
        if scene is not None and len(sh) and len(halfs):
            _, gp = sat(
                halfs[None, :, None, :],
                Rw[:, :, None],
                cw[:, :, None],
                sh[None, None],
                sR[None, None],
                sT[None, None],
            )
            g = np.minimum(g, gp.reshape(hi - lo, -1).min(axis=1))
        if spec.check_self:
            for gi, gj in spec.self_pairs:
                for b1 in [k for k, gk in enumerate(groups) if gk == gi]:
                    for b2 in [k for k, gk in enumerate(groups) if gk == gj]:
                        _, gp = sat(
                            halfs[b1], Rw[:, b1], cw[:, b1], halfs[b2], Rw[:, b2], cw[:, b2]
                        )
                        g = np.minimum(g, gp)
        gmin[lo:hi] = g
    return gmin


def is_in_collision(spec: RobotSpec, scene, q, margin=None, chunk=20000):
    """Boolean (N,) [and, with ``margin``, the near-margin mask]."""
    g = collision_margins(spec, scene, q, chunk)
    hit = ~(g > 0)
    if margin is None:
        return hit
    return hit, np.abs(g) <= margin


def is_in_self_collision(spec: RobotSpec, q):
    """robot.py:239-242 (ignores ``do_check_self_collision``)."""
    s2 = RobotSpec(**{**spec.__dict__, "check_self": True})
    return ~(collision_margins(s2, None, q) > 0)


# ----------------------------------------------------------------------------
# discretised path sweep
# ----------------------------------------------------------------------------
def interpolation_path(q_start, q_goal, max_q_step):
    """robot.py:229-237: ``ceil(|dq|_2 / step)`` points of ``linspace(0, 1, n)``."""
    q_start, q_goal = np.asarray(q_start, float), np.asarray(q_goal, float)
    n = int(np.ceil(np.linalg.norm(q_goal - q_start) / max_q_step))
    S = np.linspace(0, 1, n)
    return np.array([(1 - s) * q_start + s * q_goal for s in S]).reshape(n, q_start.shape[0])


def is_path_in_collision_discrete(spec, scene, q_start, q_goal, max_q_step=0.1):
    """robot.py:275-283 for a batch of segments (E, ndof) x2 -> bool (E,)."""
    q_start = np.atleast_2d(q_start)
    q_goal = np.atleast_2d(q_goal)
    out = np.zeros(len(q_start), dtype=bool)
    for e in range(len(q_start)):
        P = interpolation_path(q_start[e], q_goal[e], max_q_step)
        out[e] = bool(len(P)) and bool(is_in_collision(spec, scene, P).any())
    return out


# ----------------------------------------------------------------------------
# swept collision (fcl.continuousCollide as the reference calls it; oracle/fcl_ccd.py)
# ----------------------------------------------------------------------------
CCD_SAMPLES = 10  # FCL: min(num_max_iterations = 10, ceil(1 / toc_err))


def _mat_to_quat(m):
    """Eigen ``Quaternion(Matrix3)`` for a stack (...,3,3) -> (...,4) as (w, x, y, z)."""
    m = np.asarray(m, dtype=np.float64)
    t = m[..., 0, 0] + m[..., 1, 1] + m[..., 2, 2]
    out = np.empty(m.shape[:-2] + (4,))
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.sqrt(np.maximum(t + 1.0, 0.0))
        w0 = 0.5 * s
        r = 0.5 / s
        cand = [
            np.stack(
                [w0, (m[..., 2, 1] - m[..., 1, 2]) * r, (m[..., 0, 2] - m[..., 2, 0]) * r,
                 (m[..., 1, 0] - m[..., 0, 1]) * r], axis=-1)
        ]
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(np.maximum(m[..., i, i] - m[..., j, j] - m[..., k, k] + 1.0, 0.0))
            r = 0.5 / s
            c = np.empty(m.shape[:-2] + (4,))
            c[..., 1 + i] = 0.5 * s
            c[..., 0] = (m[..., k, j] - m[..., j, k]) * r
            c[..., 1 + j] = (m[..., j, i] + m[..., i, j]) * r
            c[..., 1 + k] = (m[..., k, i] + m[..., i, k]) * r
            cand.append(c)
    i_max = np.zeros(t.shape, dtype=int)
    i_max = np.where(m[..., 1, 1] > m[..., 0, 0], 1, i_max)
    dia = np.take_along_axis(
        np.stack([m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]], axis=-1), i_max[..., None], axis=-1
    )[..., 0]
    i_max = np.where(m[..., 2, 2] > dia, 2, i_max)
    sel = np.where(t > 0.0, 0, 1 + i_max)
    for b in range(4):
        out = np.where((sel == b)[..., None], cand[b], out)
    return out


def _quat_to_mat(q):
    w, x, y, z = (q[..., i] for i in range(4))
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz)
    R[..., 0, 1] = txy - twz
    R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz
    R[..., 1, 1] = 1 - (txx + tzz)
    R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy
    R[..., 2, 1] = tyz + twx
    R[..., 2, 2] = 1 - (txx + tyy)
    return R


def _quat_mul(a, b):
    aw, ax, ay, az = (a[..., i] for i in range(4))
    bw, bx, by, bz = (b[..., i] for i in range(4))
    return np.stack(
        [aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
         aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx], axis=-1)


def interp_motion(R1, t1, R2, t2, n_samples=CCD_SAMPLES):
    """FCL ``InterpMotion`` sampled at ``t = i / (n - 1)``: stacks (..., n, 3, 3), (..., n, 3)."""
    q = _mat_to_quat(R2 @ np.swapaxes(R1, -1, -2))
    n = np.linalg.norm(q[..., 1:], axis=-1)
    with np.errstate(invalid="ignore", divide="ignore"):
        angle = np.where(n != 0.0, 2.0 * np.arctan2(n, np.abs(q[..., 0])), 0.0)
        axis = q[..., 1:] / np.where(q[..., 0] < 0.0, -n, n)[..., None]
    axis = np.where((n != 0.0)[..., None], axis, np.array([1.0, 0.0, 0.0]))
    q1 = _mat_to_quat(R1)
    ts = np.arange(n_samples) / (n_samples - 1)
    half = 0.5 * angle[..., None] * ts
    dq = np.concatenate([np.cos(half)[..., None], np.sin(half)[..., None] * axis[..., None, :]], axis=-1)
    R = _quat_to_mat(_quat_mul(dq, q1[..., None, :]))
    p = t1[..., None, :] + ts[:, None] * (t2 - t1)[..., None, :]
    return R, p


def swept_margins(spec: RobotSpec, scene, q_start, q_goal, chunk=5000, n_samples=CCD_SAMPLES):
    """Per (q_start, q_goal) pair the smallest SAT decision quantity over everything the
    stateless ``Robot.is_path_in_collision`` (robot.py:285-317) tests: tool and link boxes
    as moving objects between their start and target world poses (geometry.py:68-81 ->
    shapes.py:60-75 -> FCL naive solver over ``InterpMotion``), base boxes fixed at
    ``tf_base``.  No self collision (robot.py:288).  ``g <= 0`` <=> the reference returns True
    on fresh ``Shape`` objects."""
    q_start = np.atleast_2d(np.asarray(q_start, dtype=np.float64))
    q_goal = np.atleast_2d(np.asarray(q_goal, dtype=np.float64))
    N = q_start.shape[0]
    sh, sR, sT = scene if isinstance(scene, tuple) else scene_arrays(scene)
    gmin = np.full(N, np.inf)
    if len(sh) == 0 or not spec.boxes:
        return gmin
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        Ra, ca = world_boxes(spec, q_start[lo:hi])
        Rb, cb = world_boxes(spec, q_goal[lo:hi])
        g = np.full(hi - lo, np.inf)
        for b, box in enumerate(spec.boxes):
            if box.group == GROUP_BASE:
                R, p = Ra[:, b, None], ca[:, b, None]  # robot.py:303-306: fixed
            else:
                R, p = interp_motion(Ra[:, b], ca[:, b], Rb[:, b], cb[:, b], n_samples)
            _, gp = sat(box.half, R[:, :, None], p[:, :, None], sh[None, None], sR[None, None],
                        sT[None, None])
            g = np.minimum(g, gp.reshape(hi - lo, -1).min(axis=1))
        gmin[lo:hi] = g
    return gmin


def is_path_in_collision(spec: RobotSpec, scene, q_start, q_goal, margin=None):
    g = swept_margins(spec, scene, q_start, q_goal)
    hit = ~(g > 0)
    if margin is None:
        return hit
    return hit, np.abs(g) <= margin


# ----------------------------------------------------------------------------
# analytical inverse kinematics
# ----------------------------------------------------------------------------
def tf_inverse(T):
    """robot_examples.py:14-23."""
    Ti = np.zeros_like(T)
    Rt = np.swapaxes(T[..., :3, :3], -1, -2)
    Ti[..., :3, :3] = Rt
    Ti[..., :3, 3] = np.einsum("...ij,...j->...i", -Rt, T[..., :3, 3])
    Ti[..., 3, 3] = 1.0
    return Ti


def _c3_branch(c3, tol):
    """Reach test + snap of arm_2.py:37-45 / anthropomorphic_arm.py:16-20."""
    reachable = ~((c3 > (1 + tol)) | (c3 < -(1 + tol)))
    snap = reachable & ((c3 > (1 - tol)) | (c3 < -(1 - tol)))
    c3 = np.where(snap, np.sign(c3), c3)
    with np.errstate(invalid="ignore"):
        s3 = np.sqrt(1 - c3**2)
    return reachable, c3, s3


def ik_arm2(p, a1, a2, a3):
    """inverse_kinematics/arm_2.py:5-104.  ``p`` (N,3) -> (sol (N,4,3), valid (N,4)),
    slots ordered [pos_a, pos_b, neg_a, neg_b]; unreachable slots are NaN."""
    p = np.atleast_2d(p)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    tol = 1e-6
    q1_pos = np.arctan2(y, x)
    q1_neg = np.arctan2(-y, -x)
    den = 2 * a2 * a3
    num = a1**2 - a2**2 - a3**2 + x**2 + y**2 + z**2
    c1, s1 = np.cos(q1_pos), np.sin(q1_pos)
    r_pos, c3, s3 = _c3_branch((num - 2 * a1 * s1 * y - 2 * a1 * x * c1) / den, tol)
    q3_pos_a, q3_pos_b = np.arctan2(s3, c3), np.arctan2(-s3, c3)
    c1, s1 = np.cos(q1_neg), np.sin(q1_neg)
    r_neg, c3, s3 = _c3_branch((num - 2 * a1 * s1 * y - 2 * a1 * x * c1) / den, tol)
    q3_neg_a, q3_neg_b = np.arctan2(s3, c3), np.arctan2(-s3, c3)
    L = np.sqrt(x**2 + y**2)
    c3, s3 = np.cos(q3_pos_b), np.sin(q3_pos_a)  # sic, arm_2.py:65
    q2_pos_a = np.arctan2(
        -a3 * s3 * (L - a1) + z * (a2 + a3 * c3), a3 * s3 * z + (L - a1) * (a2 + a3 * c3)
    )
    q2_pos_b = np.arctan2(
        a3 * s3 * (L - a1) + z * (a2 + a3 * c3), -a3 * s3 * z + (L - a1) * (a2 + a3 * c3)
    )
    s3, c3 = np.sin(q3_neg_a), np.cos(q3_neg_b)  # arm_2.py:76-77
    q2_neg_a = np.arctan2(
        a3 * s3 * (L + a1) + z * (a2 + a3 * c3), a3 * s3 * z - (L + a1) * (a2 + a3 * c3)
    )
    q2_neg_b = np.arctan2(
        -a3 * s3 * (L + a1) + z * (a2 + a3 * c3), -a3 * s3 * z - (L + a1) * (a2 + a3 * c3)
    )
    sol = np.stack(
        [
            np.stack([q1_pos, q2_pos_a, q3_pos_a], -1),
            np.stack([q1_pos, q2_pos_b, q3_pos_b], -1),
            np.stack([q1_neg, q2_neg_a, q3_neg_a], -1),
            np.stack([q1_neg, q2_neg_b, q3_neg_b], -1),
        ],
        axis=1,
    )
    valid = np.stack([r_pos, r_pos, r_neg, r_neg], axis=1)
    sol = np.where(valid[..., None], sol, np.nan)
    return sol, valid


def ik_spherical_wrist(R, R_base):
    """inverse_kinematics/spherical_wrist.py:5-30.  (N,3,3) x2 -> (N,2,3)."""
    Rrel = np.swapaxes(R_base, -1, -2) @ R
    n, s, a = Rrel[..., :, 0], Rrel[..., :, 1], Rrel[..., :, 2]
    A = np.sqrt(a[..., 0] ** 2 + a[..., 1] ** 2)
    sol1 = np.stack(
        [np.arctan2(a[..., 1], a[..., 0]), np.arctan2(A, a[..., 2]), np.arctan2(s[..., 2], -n[..., 2])],
        -1,
    )
    sol2 = np.stack(
        [
            np.arctan2(-a[..., 1], -a[..., 0]),
            np.arctan2(-A, a[..., 2]),
            np.arctan2(-s[..., 2], n[..., 2]),
        ],
        -1,
    )
    return np.stack([sol1, sol2], axis=-2)


def ik_anthropomorphic(p, l2, l3):
    """inverse_kinematics/anthropomorphic_arm.py:5-55.  (N,3) -> (sol (N,4,3), valid (N,))."""
    p = np.atleast_2d(p)
    px, py, pz = p[:, 0], p[:, 1], p[:, 2]
    tol = 1e-6
    Lps = px**2 + py**2 + pz**2
    reachable, c3, s3 = _c3_branch((Lps - l2**2 - l3**2) / (2 * l2 * l3), tol)
    t3_1, t3_2 = np.arctan2(s3, c3), np.arctan2(-s3, c3)
    Lxy = np.sqrt(px**2 + py**2)
    A = l2 + l3 * c3
    B = l3 * s3
    t2_1 = np.arctan2(pz * A - Lxy * B, Lxy * A + pz * B)
    t2_2 = np.arctan2(pz * A + Lxy * B, -Lxy * A + pz * B)
    t2_3 = np.arctan2(pz * A + Lxy * B, Lxy * A - pz * B)
    t2_4 = np.arctan2(pz * A - Lxy * B, -Lxy * A - pz * B)
    t1_1 = np.arctan2(py, px)
    t1_2 = np.arctan2(-py, -px)
    sol = np.stack(
        [
            np.stack([t1_1, t2_1, t3_1], -1),
            np.stack([t1_1, t2_3, t3_2], -1),
            np.stack([t1_2, t2_2, t3_1], -1),
            np.stack([t1_2, t2_4, t3_2], -1),
        ],
        axis=1,
    )
    sol = np.where(reachable[:, None, None], sol, np.nan)
    return sol, reachable


def arm2_spec(a1, a2, a3) -> RobotSpec:
    """Kinematics of ``Arm2`` (robot_examples.py:114-131)."""
    dh = np.array([[a1, np.pi / 2, 0, 0], [a2, 0, 0, 0], [a3, np.pi / 2, 0, 0]], dtype=np.float64)
    return RobotSpec(dh, np.zeros(3, bool), np.eye(4), None)


def ik_kuka(spec: RobotSpec, T):
    """``Kuka.ik`` (robot_examples.py:188-218).  ``T`` (N,4,4) ->
    (sol (N,8,6), valid (N,8)); slot = 2*arm_slot + wrist_slot, the reference's
    list is the valid slots in slot order.  Invalid slots are NaN."""
    T = np.asarray(T, dtype=np.float64).reshape(-1, 4, 4)
    a1, a2, d4, d6 = spec.dh[0, 0], spec.dh[1, 0], spec.dh[3, 2], spec.dh[5, 2]
    Tw = tf_inverse(spec.tf_base) @ T
    if spec.tf_tool_tip is not None:
        Tw = Tw @ tf_inverse(spec.tf_tool_tip)
    Tw = Tw.copy()
    Tw[:, :3, 3] = Tw[:, :3, 3] - Tw[:, :3, :3] @ np.array([0.0, 0.0, d6])
    arm = arm2_spec(a1, a2, d4)
    sol_arm, valid_arm = ik_arm2(Tw[:, :3, 3], a1, a2, d4)
    N = T.shape[0]
    sol = np.full((N, 8, 6), np.nan)
    valid = np.zeros((N, 8), dtype=bool)
    for k in range(4):
        q_arm = sol_arm[:, k].copy()
        q_arm[:, 2] = q_arm[:, 2] + np.pi / 2
        ok = valid_arm[:, k]
        R_bw = fk(arm, np.where(ok[:, None], q_arm, 0.0))[:, :3, :3]
        q_wrist = ik_spherical_wrist(Tw[:, :3, :3], R_bw)
        for w in range(2):
            sol[:, 2 * k + w, :3] = q_arm
            sol[:, 2 * k + w, 3:] = q_wrist[:, w]
            valid[:, 2 * k + w] = ok
    sol[~valid] = np.nan
    return sol, valid


# ----------------------------------------------------------------------------
# Jacobians (analytic restatement of the casadi AD results, robot.py:157-171)
# ----------------------------------------------------------------------------
def _fk_with_partials(spec: RobotSpec, q, use_base=False, use_tool=False):
    """End-effector T (N,4,4) and dT/dq_i (N,ndof,4,4).

    The reference bakes ``tf_base`` / ``tf_tool_tip`` into the casadi functions
    *as of construction* (identity / none, robot.py:194-195), hence the flags."""
    Ti = dh_transforms(spec, q)
    N, ndof = Ti.shape[:2]
    # derivative of the DH matrix wrt its joint variable
    a = spec.dh[:, 0]
    alpha = spec.dh[:, 1]
    theta = np.where(spec.prismatic, spec.dh[:, 3], np.atleast_2d(q))
    ct, st = np.cos(theta), np.sin(theta)
    ca, sa = np.cos(alpha), np.sin(alpha)
    dTi = np.zeros_like(Ti)
    rev = ~spec.prismatic
    dTi[..., 0, 0] = np.where(rev, -st, 0)
    dTi[..., 0, 1] = np.where(rev, -ct * ca, 0)
    dTi[..., 0, 2] = np.where(rev, ct * sa, 0)
    dTi[..., 0, 3] = np.where(rev, -a * st, 0)
    dTi[..., 1, 0] = np.where(rev, ct, 0)
    dTi[..., 1, 1] = np.where(rev, -st * ca, 0)
    dTi[..., 1, 2] = np.where(rev, st * sa, 0)
    dTi[..., 1, 3] = np.where(rev, a * ct, 0)
    dTi[..., 2, 3] = np.where(rev, 0, 1.0)
    base = spec.tf_base if use_base else np.eye(4)
    tail = spec.tf_tool_tip if (use_tool and spec.tf_tool_tip is not None) else np.eye(4)
    # prefix[i] = base @ T_0..T_{i-1};  suffix[i] = T_{i+1}..T_{n-1} @ tail
    prefix = [np.broadcast_to(base, (N, 4, 4))]
    for i in range(ndof):
        prefix.append(prefix[-1] @ Ti[:, i])
    suffix = [None] * (ndof + 1)
    suffix[ndof] = np.broadcast_to(tail, (N, 4, 4))
    for i in range(ndof - 1, -1, -1):
        suffix[i] = Ti[:, i] @ suffix[i + 1]
    T = prefix[ndof] @ tail
    dT = np.stack([prefix[i] @ dTi[:, i] @ suffix[i + 1] for i in range(ndof)], axis=1)
    return T, dT


def jacobian_position(spec: RobotSpec, q, use_base=False, use_tool=False):
    """(N, 3, ndof): d p_ee / d q; robot.py:163-166."""
    _, dT = _fk_with_partials(spec, q, use_base, use_tool)
    return np.swapaxes(dT[:, :, :3, 3], 1, 2)


def jacobian_rpy(spec: RobotSpec, q, use_base=False, use_tool=False):
    """(N, 6, ndof): Jacobian of the reference 6-vector; robot.py:168-171 with
    the angle formulas of robot.py:138-144 differentiated by the chain rule."""
    T, dT = _fk_with_partials(spec, q, use_base, use_tool)
    N, ndof = dT.shape[:2]
    J = np.empty((N, 6, ndof))
    J[:, :3] = np.swapaxes(dT[:, :, :3, 3], 1, 2)

    def datan2(y, x, dy, dx):  # d atan2(y, x)
        return (x[:, None] * dy - y[:, None] * dx) / (x * x + y * y)[:, None]

    T00, T01, T02 = T[:, 0, 0], T[:, 0, 1], T[:, 0, 2]
    T12, T22 = T[:, 1, 2], T[:, 2, 2]
    d00, d01, d02 = dT[:, :, 0, 0], dT[:, :, 0, 1], dT[:, :, 0, 2]
    d12, d22 = dT[:, :, 1, 2], dT[:, :, 2, 2]
    s = np.sqrt(T00 * T00 + T01 * T01)
    ds = (T00[:, None] * d00 + T01[:, None] * d01) / s[:, None]
    J[:, 3] = datan2(-T12, T22, -d12, d22)
    J[:, 4] = datan2(T02, s, d02, ds)
    J[:, 5] = datan2(-T01, T22, -d01, d22)
    return J
