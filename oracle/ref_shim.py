"""Import the *verbatim* reference (``/root/reference/src/acrobotics``) with stub
modules standing in for its absent third-party dependencies (TEST
INFRASTRUCTURE; only usable where ``/root/reference`` exists, i.e. in the build
container - never on the GPU box).

What runs unchanged: the reference's own ``link.py``, ``robot.py``,
``geometry.py``, ``shapes.py``, ``robot_examples.py`` and
``inverse_kinematics/*``.  What is restated because the dependency is missing:

* ``fcl.collide`` for Box-Box -> ``oracle.fcl_boxbox`` (FCL ``boxBox2``);
* ``fcl.continuousCollide`` (CCDM_LINEAR, naive solver, python-fcl's sticky result) ->
  ``oracle.fcl_ccd``;
* ``acrolib.geometry.translation / pose_x / pose_z / rot_x / rot_y / rot_z``
  (textbook homogeneous transforms; semantics confirmed by reference
  ``tests/test_robot_kinematics.py:100-107`` and the README golden vectors);
* casadi Jacobian builders are replaced by no-ops (``robot.py:163-171``), so
  ``jacobian_position`` / ``jacobian_rpy`` are NOT available through the shim.

Usage::

    from oracle.ref_shim import load_reference
    ab = load_reference()          # the reference package, e.g. ab.Kuka()
"""
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REFERENCE_SRC = "/root/reference/src"
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "acrobotics"))


# -- acrolib.geometry helpers that touch the hot path -------------------------
def translation(x, y, z):
    tf = np.eye(4)
    tf[:3, 3] = [x, y, z]
    return tf


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=float)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=float)


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=float)


def _pose(rot, a, x, y, z):
    tf = np.eye(4)
    tf[:3, :3] = rot(a)
    tf[:3, 3] = [x, y, z]
    return tf


def pose_x(a, x, y, z):
    return _pose(rot_x, a, x, y, z)


def pose_y(a, x, y, z):
    return _pose(rot_y, a, x, y, z)


def pose_z(a, x, y, z):
    return _pose(rot_z, a, x, y, z)


# -- fcl stub -------------------------------------------------------------------
class _FclBox:
    def __init__(self, x, y, z):
        self.side = np.array([x, y, z], dtype=np.float64)


class _FclTransform:
    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.array(R, dtype=np.float64)
        self.t = np.zeros(3) if t is None else np.array(t, dtype=np.float64)


class _FclCollisionObject:
    def __init__(self, geom, tf=None):
        self.geom = geom
        self.tf = tf if tf is not None else _FclTransform()


def _fcl_collide(o1, o2, request=None, result=None):
    from .fcl_boxbox import box_box

    hit, _ = box_box(o1.geom.side, o1.tf.R, o1.tf.t, o2.geom.side, o2.tf.R, o2.tf.t)
    return int(hit)


def _install_stubs():
    from . import fcl_ccd

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod(
        "fcl",
        Box=_FclBox,
        Cylinder=MagicMock(),
        Transform=_FclTransform,
        CollisionObject=_FclCollisionObject,
        collide=_fcl_collide,
        continuousCollide=fcl_ccd.continuous_collide,
        CollisionGeometry=MagicMock(),
        CollisionRequest=MagicMock,
        CollisionResult=MagicMock,
        ContinuousCollisionRequest=fcl_ccd.ContinuousCollisionRequest,
        ContinuousCollisionResult=fcl_ccd.ContinuousCollisionResult,
        CCDMotionType=fcl_ccd.CCDMotionType,
    )
    sys.modules["casadi"] = MagicMock(name="casadi")
    mpl = mod("matplotlib")
    mpl.animation = mod("matplotlib.animation", FuncAnimation=MagicMock())
    mpl.pyplot = mod("matplotlib.pyplot")
    mpl.pyplot.__dict__.update(figure=MagicMock(), show=MagicMock())
    acrolib = mod("acrolib")
    acrolib.__path__ = []
    mod("acrolib.plotting", plot_reference_frame=MagicMock(), get_default_axes3d=MagicMock())
    mod(
        "acrolib.geometry",
        translation=translation,
        pose_x=pose_x,
        pose_y=pose_y,
        pose_z=pose_z,
        rot_x=rot_x,
        rot_y=rot_y,
        rot_z=rot_z,
        rotation_matrix_to_rpy=MagicMock(),
        rpy_to_rot_mat=MagicMock(),
        tf_inverse=MagicMock(),
        quat_distance=MagicMock(),
    )
    mod("acrolib.sampling", Sampler=MagicMock(), SampleMethod=MagicMock())
    mod("acrolib.quaternion", Quaternion=MagicMock())
    mod(
        "acrolib.dynamic_programming",
        shortest_path=MagicMock(),
        shortest_path_with_state_cost=MagicMock(),
    )
    mod(
        "acrolib.cost_functions",
        norm_l1=MagicMock(),
        norm_l2=MagicMock(),
        sum_squared=MagicMock(),
        weighted_sum_squared=MagicMock(),
    )
    mod("pyquaternion", Quaternion=MagicMock())
    mod("urdfpy", URDF=MagicMock())


def load_reference():
    """Return the reference's ``acrobotics`` package (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_SRC}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    _install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # Robot.__init__ (robot.py:194-195) builds casadi functions; neutralise the
    # builders before any Robot is constructed (and before tool_examples runs).
    import acrobotics.link  # noqa: F401
    import acrobotics.robot as ref_robot

    ref_robot.RobotCasadiKinematics._create_jacobian_position = lambda self: None
    ref_robot.RobotCasadiKinematics._create_jacobian_rpy = lambda self: None
    import acrobotics as ab

    _loaded = ab
    return ab
