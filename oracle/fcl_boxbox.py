"""Scalar FP64 restatement of FCL's Box-Box narrow phase (TEST INFRASTRUCTURE).

The reference calls ``fcl.collide(o1, o2, request, result)`` on two
``fcl.CollisionObject(fcl.Box(dx, dy, dz), fcl.Transform(R, t))``
(reference ``src/acrobotics/shapes.py:50-58``, ``:105-114``).  python-fcl is an
un-vendored, unpinned dependency (``requirements.txt:6``); 2020-era wheels are
python-fcl 0.0.12 (FCL 0.5.0), later ones bundle FCL 0.6/0.7.  For two boxes
every one of those versions dispatches to ``boxBoxIntersect`` -> ``boxBox2``
(FCL ``narrowphase/detail/primitive_shape_algorithm/box_box-inl.h``, derived
from ODE's ``dBoxBox2``) and reports a collision iff ``return_code != 0``, i.e.
iff none of 15 candidate axes separates the boxes:

* ``p = T2 - T1``, ``pp = R1^T p``, ``A = side1/2``, ``B = side2/2``,
  ``R = R1^T R2``, ``Q = |R|``;
* face axes of box 1:  ``|pp_i| - (A_i + sum_j Q_ij B_j) > 0``  => separated;
* face axes of box 2:  ``|(R2^T p)_j| - (sum_i Q_ij A_i + B_j) > 0`` => separated;
* then ``Q += 1e-6`` (ODE's ``fudge2``) and nine edge x edge axes
  ``|pp_{i+2} R_{i+1,j} - pp_{i+1} R_{i+2,j}|
     - (A_{i+1} Q_{i+2,j} + A_{i+2} Q_{i+1,j} + B_{j+1} Q_{i,j+2} + B_{j+2} Q_{i,j+1}) > eps``
  => separated (``eps = DBL_EPSILON``; face tests use ``> 0``);
* otherwise the boxes collide (touching counts as colliding).

``box_box`` returns the boolean together with ``g``, the largest decision
quantity over the 15 axes (edge axes un-normalised, exactly the number FCL
compares).  ``g > margin`` means robustly separated, ``g < -margin`` robustly
overlapping on every axis; anything else is the "near-margin" bucket that the
parity reports list separately.
"""
import numpy as np

FUDGE2 = 1.0e-6
EPS = float(np.finfo(np.float64).eps)


def box_box(side1, R1, T1, side2, R2, T2):
    """Return ``(collide, g)`` for two oriented boxes (full side lengths)."""
    R1 = np.asarray(R1, dtype=np.float64)
    R2 = np.asarray(R2, dtype=np.float64)
    p = np.asarray(T2, dtype=np.float64) - np.asarray(T1, dtype=np.float64)
    pp = R1.T @ p
    A = 0.5 * np.asarray(side1, dtype=np.float64)
    B = 0.5 * np.asarray(side2, dtype=np.float64)
    R = R1.T @ R2
    Q = np.abs(R)

    collide = True
    g = -np.inf
    # face axes of box 1 (u1, u2, u3)
    for i in range(3):
        s2 = abs(pp[i]) - (Q[i, 0] * B[0] + Q[i, 1] * B[1] + Q[i, 2] * B[2] + A[i])
        g = max(g, s2)
        if s2 > 0:
            collide = False
    # face axes of box 2 (v1, v2, v3)
    for j in range(3):
        tmp = R2[0, j] * p[0] + R2[1, j] * p[1] + R2[2, j] * p[2]
        s2 = abs(tmp) - (Q[0, j] * A[0] + Q[1, j] * A[1] + Q[2, j] * A[2] + B[j])
        g = max(g, s2)
        if s2 > 0:
            collide = False
    # edge x edge axes, with the fudge added to |R|
    Qf = Q + FUDGE2
    for i in range(3):
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            tmp = pp[i2] * R[i1, j] - pp[i1] * R[i2, j]
            s2 = abs(tmp) - (
                A[i1] * Qf[i2, j] + A[i2] * Qf[i1, j] + B[j1] * Qf[i, j2] + B[j2] * Qf[i, j1]
            )
            g = max(g, s2 - EPS)
            if s2 > EPS:
                collide = False
    return collide, float(g)


def collide_tf(side1, tf1, side2, tf2):
    """Boolean only, with 4x4 homogeneous transforms (the ``fcl.collide`` seam)."""
    c, _ = box_box(side1, tf1[:3, :3], tf1[:3, 3], side2, tf2[:3, :3], tf2[:3, 3])
    return c
