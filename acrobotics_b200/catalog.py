"""Ready-made robots: the reference's ``robot_examples`` (``Kuka``
``src/acrobotics/robot_examples.py:137-218``, ``KukaOnRail`` ``:221-296``,
``Arm2`` ``:114-134``, ``AnthropomorphicArm`` ``:93-111``, ``SphericalWrist``
``:69-90``, ``SphericalArm`` ``:51-66``, ``PlanarArm`` ``:26-48``) as data tables.

``Kuka.ik`` / ``Kuka.ik_batch`` run the batched analytical-IK kernel
(``csrc/jac_ik.cu``); ``Kuka.ik_filtered_batch`` fuses IK with the collision
filter of the sampling-based planner (``path/path_pt_base.py:76-85``, ``:107-109``).
"""
import numpy as np
import torch

from . import _native, closed_form_ik, device
from .boxes import Box, Scene
from .chain import DHLink, JointType, Link
from .model import IKResult, JointLimit, Robot
from .transforms import tf_inverse, translation

PI = np.pi
R, P = JointType.revolute, JointType.prismatic


def _build(rows):
    """rows: (a, alpha, d, theta, joint type, box sides, box offset xyz or None)."""
    links = []
    for a, alpha, d, theta, jt, sides, off in rows:
        tf = np.eye(4) if off is None else translation(*off)
        links.append(Link(DHLink(a, alpha, d, theta), jt, Scene([Box(*sides)], [tf])))
    return links


class PlanarArm(Robot):
    """Three parallel revolute joints (Siciliano p. 69)."""

    def __init__(self, a1=1, a2=1, a3=1):
        super().__init__(
            _build([(a, 0, 0, 0, R, (a, 0.1, 0.1), (-a / 2, 0, 0)) for a in (a1, a2, a3)])
        )

    def ik(self, T):
        raise NotImplementedError("planar_arm.ik is outside the GPU hot path (see DESIGN.md)")


class SphericalArm(Robot):
    """Two revolute joints and a prismatic one (Siciliano p. 72)."""

    def __init__(self, d2=1):
        bar = (1, 0.1, 0.1)
        super().__init__(
            _build(
                [
                    (0, -PI / 2, 0, 0, R, bar, None),
                    (0, PI / 2, d2, 0, R, bar, None),
                    (0, 0, 0, 0, P, bar, None),
                ]
            )
        )


class SphericalWrist(Robot):
    """Three intersecting revolute axes (Siciliano p. 75)."""

    def __init__(self, d3=1):
        bar = (1, 0.1, 0.1)
        super().__init__(
            _build(
                [
                    (0, -PI / 2, 0, 0, R, bar, None),
                    (0, PI / 2, 0, 0, R, bar, None),
                    (0, 0, d3, 0, R, bar, None),
                ]
            )
        )

    def ik(self, T):
        return closed_form_ik.spherical_wrist.ik(T, self.tf_base)


class AnthropomorphicArm(Robot):
    """Shoulder + elbow arm without offsets (Siciliano p. 73)."""

    def __init__(self, a2=1, a3=1):
        super().__init__(
            _build(
                [
                    (0, PI / 2, 0, 0, R, (0.3, 0.3, 0.1), (0, 0, 0.05)),
                    (a2, 0, 0, 0, R, (a2, 0.1, 0.1), (-a2 / 2, 0, 0)),
                    (a3, 0, 0, 0, R, (a3, 0.1, 0.1), (-a3 / 2, 0, 0)),
                ]
            )
        )

    def ik(self, T):
        return closed_form_ik.anthropomorphic_arm.ik(T, self.links)


class Arm2(Robot):
    """Articulated arm with a shoulder offset ``a1``; the last frame is turned so
    that z points along a would-be fourth joint."""

    def __init__(self, a1=1, a2=1, a3=1):
        super().__init__(
            _build(
                [
                    (a1, PI / 2, 0, 0, R, (a1, 0.1, 0.1), None),
                    (a2, 0, 0, 0, R, (a2, 0.1, 0.1), None),
                    (a3, PI / 2, 0, 0, R, (a3, 0.1, 0.1), None),
                ]
            )
        )

    def ik(self, T):
        return closed_form_ik.arm_2.ik(T, self.links)


# Kuka KR5-like arm: box sides and box centre offsets per link
_KUKA_BOXES = [
    ((0.3, 0.2, 0.1), (-0.09, 0, 0.05)),
    ((0.8, 0.2, 0.1), (-0.3, 0, -0.05)),
    ((0.2, 0.1, 0.5), (0, 0.05, 0.17)),
    ((0.1, 0.2, 0.1), (0, 0.1, 0)),
    ((0.1, 0.1, 0.085), (0, 0, 0.085 / 2)),
    ((0.1, 0.1, 0.03), (0, 0, -0.03 / 2)),
]


def _kuka_rows(a1, a2, d4, d6):
    dh = [
        (a1, PI / 2, 0),
        (a2, 0, 0),
        (0, PI / 2, 0),
        (0, -PI / 2, d4),
        (0, PI / 2, 0),
        (0, 0, d6),
    ]
    return [(a, al, d, 0, R, sides, off) for (a, al, d), (sides, off) in zip(dh, _KUKA_BOXES)]


class Kuka(Robot):
    """6-dof arm = Arm2 + spherical wrist, with analytical inverse kinematics."""

    def __init__(self, a1=0.18, a2=0.6, d4=0.62, d6=0.115):
        deg = np.deg2rad
        limits = [
            JointLimit(deg(-155), deg(155)),
            JointLimit(deg(-180), deg(65)),
            JointLimit(deg(-15), deg(158)),
            JointLimit(deg(-350), deg(350)),
            JointLimit(deg(-130), deg(130)),
            JointLimit(deg(-350), deg(350)),
        ]
        super().__init__(_build(_kuka_rows(a1, a2, d4, d6)), limits)
        self.arm = Arm2(a1=a1, a2=a2, a3=d4)
        self.wrist = SphericalWrist(d3=d6)

    def _poses(self, T):
        t, host = device.to_device(T, torch.float64, ndim=3)
        return t.reshape(-1, 4, 4), host

    def ik_batch(self, T):
        """T (N,4,4) -> (q (N,8,6), valid (N,8) bool).  Slot ``2*arm + wrist`` in
        the reference's order; slots that do not exist are NaN."""
        Td, host = self._poses(T)
        n = Td.shape[0]
        q = torch.empty((n, 8, 6), dtype=torch.float64, device="cuda")
        bits = torch.empty(n, dtype=torch.uint8, device="cuda")
        _native.check(
            _native.load().acro_ik_kuka_f64(
                self._handle(kinematics_only=True), Td.data_ptr(), n, q.data_ptr(),
                bits.data_ptr(), device.current_stream_ptr(),
            ),
            "acro_ik_kuka_f64",
        )
        valid = closed_form_ik._bits_to_bool(bits, 8)
        return device.to_host(q, host), device.to_host(valid, host)

    def ik_filtered_batch(self, T, scene, _force_self=False):
        """IK followed by the collision filter (``scene=None``: self collision only):
        -> (q (N,8,6), valid (N,8), collision_free (N,8))."""
        Td, host = self._poses(T)
        n = Td.shape[0]
        self.cc_checks += 8 * n
        q = torch.empty((n, 8, 6), dtype=torch.float64, device="cuda")
        bits = torch.empty(n, dtype=torch.uint8, device="cuda")
        free = torch.empty(n, dtype=torch.uint8, device="cuda")
        _native.check(
            _native.load().acro_ik_kuka_cc_f64(
                self._handle(force_self=_force_self),
                scene.native_handle() if scene is not None else None,
                Td.data_ptr(), n, q.data_ptr(), bits.data_ptr(), free.data_ptr(),
                device.current_stream_ptr(),
            ),
            "acro_ik_kuka_cc_f64",
        )
        b2b = closed_form_ik._bits_to_bool
        return (
            device.to_host(q, host),
            device.to_host(b2b(bits, 8), host),
            device.to_host(b2b(free, 8), host),
        )

    def ik(self, T) -> IKResult:
        q, valid = self.ik_batch(np.asarray(T, dtype=np.float64).reshape(1, 4, 4))
        sols = [q[0, s].copy() for s in range(8) if valid[0, s]]
        return IKResult(True, sols) if sols else IKResult(False)


class KukaOnRail(Robot):
    """Kuka on a prismatic rail: 7 joints, the first one is the rail."""

    def __init__(self, a1=0.18, a2=0.6, d4=0.62, d6=0.115):
        rail = [(0, PI / 2, 0, 0, P, (0.2, 0.2, 0.1), (0, 0, -0.15))]
        super().__init__(_build(rail + _kuka_rows(a1, a2, d4, d6)))
        self.kuka = Kuka(a1=a1, a2=a2, d4=d4, d6=d6)

    def ik_batch(self, T, q_fixed):
        """Batched ``ik`` for one rail position: T (N,4,4) -> (q (N,8,7), valid (N,8)); the
        first column of every solution is ``q_fixed``."""
        Td = np.asarray(T, dtype=np.float64).reshape(-1, 4, 4)
        Tw = tf_inverse(self.tf_base) @ Td
        if self.tf_tool_tip is not None:
            Tw = Tw @ tf_inverse(self.tf_tool_tip)
        self.kuka.tf_base = self.links[0].get_link_relative_transform(q_fixed)
        q6, valid = self.kuka.ik_batch(Tw)
        q7 = np.concatenate([np.full(q6.shape[:2] + (1,), float(q_fixed)), q6], axis=2)
        q7[~valid] = np.nan
        return q7, valid

    def ik(self, T, q_fixed) -> IKResult:
        Tw = tf_inverse(self.tf_base) @ np.asarray(T, dtype=np.float64)
        if self.tf_tool_tip is not None:
            Tw = Tw @ tf_inverse(self.tf_tool_tip)
        # the rail position fixes the base of the 6-dof helper robot
        self.kuka.tf_base = self.links[0].get_link_relative_transform(q_fixed)
        res = self.kuka.ik(Tw)
        if not res.success:
            return IKResult(False)
        return IKResult(True, [np.hstack((q_fixed, qi)) for qi in res.solutions])
