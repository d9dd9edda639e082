"""Alias of the reference package path ``acrobotics.inverse_kinematics``."""
from ..closed_form_ik import anthropomorphic_arm, arm_2, spherical_wrist  # noqa: F401
from ..model import IKResult  # noqa: F401
