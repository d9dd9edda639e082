"""Scalar FP64 restatement of ``fcl.continuousCollide`` as the reference uses it (TEST
INFRASTRUCTURE; parity unpinned against python-fcl itself, which is absent - pinned only by
the three known answers of the reference's ``tests/test_continuous_cc.py:35-83``).

Reference call site (``src/acrobotics/shapes.py:60-75``)::

    self.c_req.ccd_motion_type = fcl.CCDMotionType.CCDM_LINEAR
    fcl.continuousCollide(o1, fcl_tf_1_target, o2, fcl_tf_2, self.c_req, self.c_res)
    return self.c_res.is_collide

with ``c_req = fcl.ContinuousCollisionRequest()`` and ``c_res =
fcl.ContinuousCollisionResult()`` created ONCE per ``Shape`` (``shapes.py:113-114``).

What FCL (0.5 - 0.7, ``narrowphase/continuous_collision-inl.h``, ``math/motion/
interp_motion-inl.h``) does for that call, restated from its published sources:

* request defaults: ``num_max_iterations = 10``, ``toc_err = 1e-4``, ``gjk_solver_type =
  GST_LIBCCD``, ``ccd_solver_type = CCDC_NAIVE``; the reference only changes the motion type;
* ``CCDM_LINEAR`` -> ``InterpMotion(tf_begin, tf_end)`` for each object (``reference_p = 0``):
  ``linear_vel = t_end - t_begin``; ``AngleAxis(R_end R_begin^T)`` (through Eigen's
  matrix -> quaternion -> angle-axis, angle in [0, pi]) gives ``angular_axis`` and
  ``angular_vel``; ``integrate(t)``: ``R(t) = (Quaternion(AngleAxis(t * angular_vel,
  angular_axis)) * Quaternion(R_begin)).toRotationMatrix()``, ``p(t) = t_begin + t *
  linear_vel``.  The box centre moves along the CHORD between its two world positions - not
  along the arc the joint motion would give it - and the orientation turns about one fixed
  axis at constant rate;
* ``continuousCollideNaive``: ``n_iter = min(num_max_iterations, ceil(1 / toc_err)) = 10``
  samples ``t = i / (n_iter - 1)``, i = 0..9 (both end poses included); at each sample a
  default discrete ``collide`` of the two shapes (for two boxes: ``boxBox2``, see
  ``oracle/fcl_boxbox.py``); the first colliding sample ends the loop with ``is_collide``
  true;
* python-fcl's wrapper (``fcl.pyx``, ``continuousCollide``) copies the C++ result into the
  Python object with ``result.is_collide = result.is_collide or cresult.is_collide``: the
  Python ``ContinuousCollisionResult`` is STICKY.  Because the reference keeps one ``c_res``
  per ``Shape``, a shape that has collided once reports True for every later query
  (``StickyRobotEmulation`` below models that; ``sticky=False`` gives the stateless query).

The reference's third known answer, ``test_ccd_3``, is only reproduced through that sticky
state (the test module shares one robot over its three tests): the chord motion of the
upper-arm box passes 2 cm from the 1 cm plate.  The stateless restatement returns True, True,
False for the three set-ups; the stateful one True, True, True.  Both are recorded in
``tests/golden/ccd.npz``.
"""
import numpy as np

from .fcl_boxbox import box_box

N_ITER = 10  # min(num_max_iterations = 10, ceil(1 / toc_err = 1e-4))


def mat_to_quat(m):
    """Eigen ``Quaternion(Matrix3)`` -> (w, x, y, z)."""
    m = np.asarray(m, dtype=np.float64)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.empty(4)
    if t > 0.0:
        s = np.sqrt(t + 1.0)
        q[0] = 0.5 * s
        s = 0.5 / s
        q[1] = (m[2, 1] - m[1, 2]) * s
        q[2] = (m[0, 2] - m[2, 0]) * s
        q[3] = (m[1, 0] - m[0, 1]) * s
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * s
        s = 0.5 / s
        q[0] = (m[k, j] - m[j, k]) * s
        q[1 + j] = (m[j, i] + m[i, j]) * s
        q[1 + k] = (m[k, i] + m[i, k]) * s
    return q


def quat_to_mat(q):
    """Eigen ``Quaternion::toRotationMatrix``."""
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [1 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1 - (txx + tyy)],
        ]
    )


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
        ]
    )


def angle_axis(m):
    """Eigen ``AngleAxis(Matrix3)``: through the quaternion, angle in [0, pi]."""
    q = mat_to_quat(m)
    n = np.linalg.norm(q[1:])
    if n != 0.0:
        angle = 2.0 * np.arctan2(n, abs(q[0]))
        if q[0] < 0.0:
            n = -n
        return angle, q[1:] / n
    return 0.0, np.array([1.0, 0.0, 0.0])


class InterpMotion:
    """FCL ``InterpMotion(tf1, tf2)`` with ``reference_p = 0``."""

    def __init__(self, R1, t1, R2, t2):
        self.R1 = np.asarray(R1, dtype=np.float64)
        self.t1 = np.asarray(t1, dtype=np.float64)
        self.linear_vel = np.asarray(t2, dtype=np.float64) - self.t1
        self.angular_vel, self.angular_axis = angle_axis(np.asarray(R2, dtype=np.float64) @ self.R1.T)
        self.q1 = mat_to_quat(self.R1)

    def at(self, dt):
        dt = min(dt, 1.0)
        half = 0.5 * dt * self.angular_vel
        dq = np.concatenate([[np.cos(half)], np.sin(half) * self.angular_axis])
        return quat_to_mat(quat_mul(dq, self.q1)), self.t1 + dt * self.linear_vel


def continuous_collide_boxes(side1, R1, t1, R1_end, t1_end, side2, R2, t2, R2_end=None, t2_end=None):
    """``continuousCollideNaive`` for two boxes.  Returns ``(is_collide, time_of_contact, g)``
    with ``g`` the smallest, over the samples tested, of the pair's largest SAT decision
    quantity (``oracle/fcl_boxbox.py``)."""
    m1 = InterpMotion(R1, t1, R1_end, t1_end)
    m2 = InterpMotion(R2, t2, R2 if R2_end is None else R2_end, t2 if t2_end is None else t2_end)
    gmin = np.inf
    for i in range(N_ITER):
        t = i / (N_ITER - 1)
        Ra, pa = m1.at(t)
        Rb, pb = m2.at(t)
        hit, g = box_box(side1, Ra, pa, side2, Rb, pb)
        gmin = min(gmin, g)
        if hit:
            return True, t, gmin
    return False, 1.0, gmin


# -- the python-fcl objects the reference touches (shapes.py:60-75, :113-114) -----------
class CCDMotionType:
    CCDM_TRANS, CCDM_LINEAR, CCDM_SCREW, CCDM_SPLINE = range(4)


class ContinuousCollisionRequest:
    def __init__(self):
        self.num_max_iterations = 10
        self.toc_err = 0.0001
        self.ccd_motion_type = CCDMotionType.CCDM_TRANS


class ContinuousCollisionResult:
    def __init__(self):
        self.is_collide = False
        self.time_of_contact = 1.0


STICKY = True  # python-fcl's result accumulation; make_golden switches it off for stateless vectors


def continuous_collide(o1, tf1_end, o2, tf2_end, request=None, result=None):
    """``fcl.continuousCollide`` of the python-fcl wrapper for two box objects."""
    request = request or ContinuousCollisionRequest()
    result = result if result is not None else ContinuousCollisionResult()
    if request.ccd_motion_type != CCDMotionType.CCDM_LINEAR:
        raise NotImplementedError("only CCDM_LINEAR is restated (shapes.py:72)")
    hit, toc, _ = continuous_collide_boxes(
        o1.geom.side, o1.tf.R, o1.tf.t, tf1_end.R, tf1_end.t,
        o2.geom.side, o2.tf.R, o2.tf.t, tf2_end.R, tf2_end.t,
    )
    if STICKY:
        result.is_collide = bool(result.is_collide) or hit
        result.time_of_contact = min(toc, result.time_of_contact)
    else:
        result.is_collide = hit
        result.time_of_contact = toc
    return toc


class StickyRobotEmulation:
    """The reference's ``Robot.is_path_in_collision`` (robot.py:285-317) over a flattened
    ``np_oracle.RobotSpec``, including the state it carries between queries: one python-fcl
    ``ContinuousCollisionResult`` per shape whose ``is_collide`` only ever turns True
    (``shapes.py:74-75``, ``:113-114``) and the ``collision_priority`` list.  Usable without
    the reference tree (the CPU tests replay the reference's three known answers with it)."""

    def __init__(self, spec, sticky=True):
        from . import np_oracle as no

        self._no = no
        self.spec = spec
        self.sticky = sticky
        self.flag = [False] * len(spec.boxes)
        self.priority = list(range(spec.ndof))

    def _shape_query(self, b, Wa, Wb, scene):
        sh, sR, sT = scene
        box = self.spec.boxes[b]
        for j in range(len(sh)):  # geometry.py:77-80: first True ends the soup query
            hit, _, _ = continuous_collide_boxes(
                2 * box.half, Wa[0][b], Wa[1][b], Wb[0][b], Wb[1][b], 2 * sh[j], sR[j], sT[j])
            res = (self.flag[b] or hit) if self.sticky else hit
            self.flag[b] = res
            if res:
                return True
        return False

    def is_path_in_collision(self, q_start, q_goal, scene):
        no = self._no
        scene = scene if isinstance(scene, tuple) else no.scene_arrays(scene)
        Ra, ca = no.world_boxes(self.spec, np.atleast_2d(q_start))
        Rb, cb = no.world_boxes(self.spec, np.atleast_2d(q_goal))
        Wa, Wb = (Ra[0], ca[0]), (Rb[0], cb[0])
        groups = [b.group for b in self.spec.boxes]
        for b in [k for k, g in enumerate(groups) if g == no.GROUP_TOOL]:  # robot.py:296-300
            if self._shape_query(b, Wa, Wb, scene):
                return True
        for b in [k for k, g in enumerate(groups) if g == no.GROUP_BASE]:  # robot.py:303-306
            box = self.spec.boxes[b]
            for j in range(len(scene[0])):
                if box_box(2 * box.half, Wa[0][b], Wa[1][b], 2 * scene[0][j], scene[1][j], scene[2][j])[0]:
                    return True
        for i in list(self.priority):  # robot.py:309-316
            for b in [k for k, g in enumerate(groups) if g == i]:
                if self._shape_query(b, Wa, Wb, scene):
                    self.priority.remove(i)
                    self.priority.insert(0, i)
                    return True
        return False
