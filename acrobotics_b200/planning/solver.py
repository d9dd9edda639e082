"""``solve`` dispatcher (reference ``src/acrobotics/planning/solver.py``)."""
from .sampling_based import sampling_based_solve
from .types import SolveMethod, Solution


def solve(planning_setup, solver_settings) -> Solution:
    if solver_settings.solve_method == SolveMethod.sampling_based:
        return sampling_based_solve(planning_setup, solver_settings)
    if solver_settings.solve_method == SolveMethod.optimization_based:
        raise NotImplementedError(
            "the optimisation-based planner (casadi / IPOPT) is outside the GPU hot path")
    raise Exception(f"Solver method {solver_settings.solve_method} not implemented")
