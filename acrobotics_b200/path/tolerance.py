"""Module path of the reference (``acrobotics.path.tolerance``); see ``ranges.py``."""
from .ranges import NoTolerance, QuaternionTolerance, SymmetricTolerance, Tolerance  # noqa: F401
