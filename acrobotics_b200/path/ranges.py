"""Sampled intervals that describe how far a path point may deviate along one coordinate.

API of the reference's ``Tolerance`` family (``src/acrobotics/path/tolerance.py:4-40``):
attributes ``lower`` / ``upper`` / ``num_samples`` / ``has_tolerance`` and the methods
``discretize()`` and ``reduce_tolerance(reduction_factor, reference_value)``.  One class
carries the behaviour; the reference's three spellings are constructors of it.
"""
import numpy as np


class Tolerance:
    """Closed interval ``[lower, upper]`` sampled at ``num_samples`` equidistant values.
    A degenerate interval (``fixed=True``) stands for "no tolerance": it samples to the single
    value 0 and never shrinks."""

    __slots__ = ("lower", "upper", "num_samples", "has_tolerance")

    def __init__(self, lower: float, upper: float, num_samples: int, fixed: bool = False):
        self.lower, self.upper, self.num_samples = lower, upper, num_samples
        self.has_tolerance = not fixed

    def discretize(self) -> np.ndarray:
        if not self.has_tolerance:
            return np.zeros(1, dtype=int)
        return np.linspace(self.lower, self.upper, self.num_samples)

    def reduce_tolerance(self, reduction_factor: float, reference_value: float) -> None:
        """Shrink the interval by ``reduction_factor`` but keep ``reference_value`` inside
        (tolerance.py:17-21)."""
        if not self.has_tolerance:
            return
        shrunk = (self.lower / reduction_factor, self.upper / reduction_factor)
        keep = (min(self.lower - reference_value, 0.0), max(self.upper - reference_value, 0.0))
        self.lower, self.upper = max(shrunk[0], keep[0]), min(shrunk[1], keep[1])

    def __repr__(self):
        kind = "Tolerance" if self.has_tolerance else "NoTolerance"
        return f"{kind}({self.lower}, {self.upper}, {self.num_samples})"


class NoTolerance(Tolerance):
    __slots__ = ()

    def __init__(self):
        super().__init__(0, 0, 1, fixed=True)


class SymmetricTolerance(Tolerance):
    __slots__ = ()

    def __init__(self, dist: float, num_samples: int):
        super().__init__(-dist, dist, num_samples)


class QuaternionTolerance(Tolerance):
    """Ball of orientations: quaternion distance at most ``dist`` (tolerance.py:43-55)."""

    __slots__ = ("dist",)

    def __init__(self, quat_distance: float):
        self.dist = quat_distance
        self.lower = self.upper = self.num_samples = None
        self.has_tolerance = True

    def discretize(self):
        raise NotImplementedError  # "Future work" in the reference as well

    def reduce_tolerance(self, reduction_factor: float, reference_value: float) -> None:
        self.dist = self.dist / reduction_factor

    def __repr__(self):
        return f"QuaternionTolerance({self.dist})"
