"""Denavit-Hartenberg links: ``DHLink``, ``JointType``, ``LinkKinematics``, ``Link``.

Interface of reference ``src/acrobotics/link.py`` (``DHLink`` ``:8``,
``JointType`` ``:11-13``, ``LinkKinematics`` ``:16-58``, ``Link`` ``:80-86``).  A
link is a record; the arithmetic happens on the GPU, where the whole chain is
evaluated per configuration (``csrc/fk.cu``).  ``get_link_relative_transform``
is kept for API compatibility and runs the same kernel on a one-link chain.
"""
from collections import namedtuple
from enum import Enum

import numpy as np

DHLink = namedtuple("DHLink", ["a", "alpha", "d", "theta"])


class JointType(Enum):
    revolute = "r"
    prismatic = "p"


class LinkKinematics:
    def __init__(self, dh_parameters: DHLink, joint_type: JointType):
        self.dh = dh_parameters
        if not isinstance(joint_type, JointType):
            try:
                joint_type = JointType(joint_type)
            except ValueError:
                raise ValueError(f"Unkown JointType: {joint_type}.") from None
        self.joint_type = joint_type
        self._single = None

    def get_link_relative_transform(self, qi):
        """4x4 transform of link i relative to link i-1 for joint value ``qi``."""
        if self._single is None:
            from .model import RobotKinematics

            self._single = RobotKinematics([LinkKinematics(self.dh, self.joint_type)])
        return self._single.fk([qi])


class Link(LinkKinematics):
    def __init__(self, dh_parameters, joint_type, geometry):
        super().__init__(dh_parameters, joint_type)
        self.geometry = geometry
