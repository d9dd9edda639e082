"""Alias of the reference module path ``acrobotics.link``."""
from .chain import DHLink, JointType, Link, LinkKinematics  # noqa: F401
