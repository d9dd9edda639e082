"""Alias of the reference module path ``acrobotics.shapes`` (boxes only)."""
from .boxes import Box, Polyhedron, transform_vector  # noqa: F401
