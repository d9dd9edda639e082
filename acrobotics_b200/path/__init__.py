"""Path points and the planner's inner loop (reference ``src/acrobotics/path``).

Only what feeds the GPU hot path is mirrored: tolerances (``path/tolerance.py``), the grid
samplers of ``TolPositionPt`` / ``TolEulerPt`` (``path/path_pt.py``), ``create_grid``
(``path/util.py:23-57``) and ``PathPt.to_joint_solutions`` (``path/path_pt_base.py:87-111``),
the latter also batched over a whole path (``joint_solutions_csr``)."""
from .path_pt import (TolEulerPt, TolPositionPt, joint_solutions_csr, path_grid_arrays,
                      path_to_joint_solutions)
from .ranges import NoTolerance, SymmetricTolerance, Tolerance
from .util import create_grid, rpy_to_rot_mat

__all__ = [
    "NoTolerance", "SymmetricTolerance", "Tolerance", "TolEulerPt", "TolPositionPt",
    "create_grid", "joint_solutions_csr", "path_grid_arrays", "path_to_joint_solutions",
    "rpy_to_rot_mat",
]
