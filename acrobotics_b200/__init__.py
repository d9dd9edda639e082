"""acrobotics_b200: the acrobotics FK + collision-checking hot path on B200.

Drop-in names of the reference's façade (``src/acrobotics/__init__.py:26-54``)
for the hot path: ``Kuka``, ``KukaOnRail``, ``Arm2``, ``AnthropomorphicArm``,
``SphericalWrist``, ``SphericalArm``, ``PlanarArm``, ``Box``, ``Scene``,
``Tool``.  Module paths mirror the reference as well (``acrobotics_b200.robot``,
``.link``, ``.geometry``, ``.shapes``, ``.robot_examples``,
``.inverse_kinematics``).  Everything computes on the GPU through
``libacro_b200.so``; importing the package does not need a GPU, calling it does.
"""
from .boxes import Box, Scene
from .catalog import (
    AnthropomorphicArm,
    Arm2,
    Kuka,
    KukaOnRail,
    PlanarArm,
    SphericalArm,
    SphericalWrist,
)
from .chain import DHLink, JointType, Link
from .model import IKResult, JointLimit, Robot, RobotKinematics, Tool

__all__ = [
    "AnthropomorphicArm", "Arm2", "Box", "DHLink", "IKResult", "JointLimit", "JointType",
    "Kuka", "KukaOnRail", "Link", "PlanarArm", "Robot", "RobotKinematics", "Scene",
    "SphericalArm", "SphericalWrist", "Tool",
]
