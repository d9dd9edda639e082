"""Module path of the reference (``acrobotics.path.tolerance``); see ``ranges.py``."""
from .ranges import NoTolerance, SymmetricTolerance, Tolerance  # noqa: F401
