"""Shared builders for the parity tests: the same robots / scenes for the
product package (GPU) and, as plain specs, for the oracle."""
import numpy as np

import acrobotics_b200 as ab
from acrobotics_b200.synthetic import scene16
from oracle import np_oracle as no


def product_scene(sides, tfs):
    return ab.Scene([ab.Box(*s) for s in sides], [np.array(t) for t in tfs])


def kuka_with_tool(g):
    """The Kuka + torch2 tool + base geometry + moved base of cc_kuka.npz."""
    k = ab.Kuka()
    k.tool = ab.Tool([ab.Box(*s) for s in g["tool_sides"]], [t for t in g["tool_tfs"]], g["tool_tip"])
    k.tf_base = g["tool_base_tf"]
    k.geometry_base = product_scene(g["tool_base_sides"], g["tool_base_tfs"])
    return k


def spec(robot):
    return no.spec_from_robot(robot)


ROBOTS = ["Kuka", "KukaOnRail", "PlanarArm", "SphericalArm", "SphericalWrist",
          "AnthropomorphicArm", "Arm2"]
