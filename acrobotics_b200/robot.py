"""Alias of the reference module path ``acrobotics.robot``."""
from .model import IKResult, JointLimit, Robot, RobotKinematics, Tool  # noqa: F401
