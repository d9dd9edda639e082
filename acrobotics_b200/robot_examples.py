"""Alias of the reference module path ``acrobotics.robot_examples``."""
from .catalog import (  # noqa: F401
    AnthropomorphicArm,
    Arm2,
    Kuka,
    KukaOnRail,
    PlanarArm,
    SphericalArm,
    SphericalWrist,
)
from .transforms import tf_inverse  # noqa: F401
