"""Alias of the reference module path ``acrobotics.geometry``."""
from .boxes import Scene  # noqa: F401
