"""Path points and the planner's inner loop (reference ``src/acrobotics/path``).

Mirrors what feeds the GPU hot path: tolerances (``path/tolerance.py``), the samplers of the
path points (``path/path_pt.py``: grid and incremental), the search strategies and settings
(``path/sampling.py``), the path factories (``path/factory.py``), ``create_grid``
(``path/util.py:23-57``) and ``PathPt.to_joint_solutions`` (``path/path_pt_base.py:87-129``),
the latter also batched over a whole path (``joint_solutions_csr``)."""
from .factory import create_arc, create_circle, create_line
from .path_pt import (FreeOrientationPt, TolEulerPt, TolPositionPt, TolQuatPt,
                      joint_solutions_csr, path_grid_arrays, path_to_joint_solutions)
from .ranges import NoTolerance, QuaternionTolerance, SymmetricTolerance, Tolerance
from .sampling import SampleMethod, Sampler, SamplingSetting, SearchStrategy
from .util import create_grid, rotation_matrix_to_rpy, rpy_to_rot_mat

__all__ = [
    "FreeOrientationPt", "NoTolerance", "QuaternionTolerance", "SampleMethod", "Sampler",
    "SamplingSetting", "SearchStrategy", "SymmetricTolerance", "Tolerance", "TolEulerPt",
    "TolPositionPt", "TolQuatPt", "create_arc", "create_circle", "create_grid", "create_line",
    "joint_solutions_csr", "path_grid_arrays", "path_to_joint_solutions",
    "rotation_matrix_to_rpy", "rpy_to_rot_mat",
]
