"""FCL-independent ground truth for Box-Box intersection (TEST INFRASTRUCTURE).

The reference describes a box as a polyhedron ``A x <= b`` (``Box.get_normals`` /
``Box.get_polyhedron``, reference ``src/acrobotics/shapes.py:150-172``): the six rows of
``A`` are ``R @ (+-e_k)`` and ``b = 0.5 * [dx, dx, dy, dy, dz, dz] + A @ t``.  Two boxes
intersect iff the stacked 12-row system is feasible.  ``depth`` solves the linear programme

    maximise  s   subject to   A1 x + s <= b1,   A2 x + s <= b2

(rows of ``A`` are unit vectors, so ``s`` is a length): ``s > 0`` - the boxes share an
interior point that is ``s`` inside every face; ``s < 0`` - every point violates some face of
one of the boxes by at least ``-s``, i.e. the boxes are separated (by at least ``-s`` and at
most ``-2 sqrt(3) s``).  This is independent of the separating-axis formulation, of the axis
order, of ODE's fudge factor and of FCL's ``> eps`` convention, which is why it is used to pin
the restated ``boxBox2`` outside a small band around contact.

``polyhedron`` is checked against the verbatim reference in ``oracle/make_golden.py``
(fixture ``tests/golden/polyhedron.npz``).
"""
import numpy as np
from scipy.optimize import linprog

_SIGNS = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)


def polyhedron(sides, tf):
    """(A (6,3), b (6,)) of a box with full side lengths ``sides`` at the 4x4 pose ``tf``
    (reference shapes.py:150-172, same row order)."""
    tf = np.asarray(tf, dtype=np.float64)
    A = _SIGNS @ tf[:3, :3].T  # row k = R @ sign_k
    dx, dy, dz = (float(v) for v in sides)
    b = 0.5 * np.array([dx, dx, dy, dy, dz, dz]) + A @ tf[:3, 3]
    return A, b


def depth(sides1, tf1, sides2, tf2):
    """Signed common depth ``s`` of two boxes (see the module docstring)."""
    A1, b1 = polyhedron(sides1, tf1)
    A2, b2 = polyhedron(sides2, tf2)
    A = np.hstack([np.vstack([A1, A2]), np.ones((12, 1))])
    b = np.concatenate([b1, b2])
    res = linprog([0.0, 0.0, 0.0, -1.0], A_ub=A, b_ub=b, bounds=[(None, None)] * 4, method="highs")
    if res.status != 0:
        raise RuntimeError(f"LP failed: {res.message}")
    return float(-res.fun)


def intersects(sides1, tf1, sides2, tf2):
    return depth(sides1, tf1, sides2, tf2) >= 0.0
