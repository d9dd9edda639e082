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
# path construction and planner settings (reference __init__.py:41-54)
from .path import (FreeOrientationPt, NoTolerance, QuaternionTolerance, SampleMethod,
                   SamplingSetting, SearchStrategy, SymmetricTolerance, Tolerance, TolEulerPt,
                   TolPositionPt, TolQuatPt, create_arc, create_circle, create_line)
from .planning import (CostFuntionType, OptSettings, PlanningSetup, SolveMethod, SolverSettings,
                       solve)

__all__ = [
    "AnthropomorphicArm", "Arm2", "Box", "CostFuntionType", "DHLink", "FreeOrientationPt",
    "IKResult", "JointLimit", "JointType", "Kuka", "KukaOnRail", "Link", "NoTolerance",
    "OptSettings", "PlanarArm", "PlanningSetup", "QuaternionTolerance", "Robot",
    "RobotKinematics", "SampleMethod", "SamplingSetting", "Scene", "SearchStrategy",
    "SolveMethod", "SolverSettings", "SphericalArm", "SphericalWrist", "SymmetricTolerance",
    "Tolerance", "TolEulerPt", "TolPositionPt", "TolQuatPt", "Tool", "create_arc",
    "create_circle", "create_line", "solve",
]
