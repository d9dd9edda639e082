"""Sampling-based planner on the GPU hot path (reference ``src/acrobotics/planning``): joint
solutions per path point, the shortest path through the resulting layered graph, and the outer
loop with shrinking tolerances.  The optimisation-based planner is out of scope."""
from .sampling_based import (COST_FUNCTIONS, calc_state_cost, create_new_pt,
                             create_reduced_tolerance_path, find_shortest_joint_path,
                             path_to_joint_solutions, sampling_based_solve, shortest_path,
                             shortest_path_csr, solver_step)
from .settings import OptSettings, SolverSettings
from .solver import solve
from .types import CostFuntionType, JointPath, PlanningSetup, Solution, SolveMethod

__all__ = ["COST_FUNCTIONS", "CostFuntionType", "JointPath", "OptSettings", "PlanningSetup",
           "Solution", "SolveMethod", "SolverSettings", "calc_state_cost", "create_new_pt",
           "create_reduced_tolerance_path", "find_shortest_joint_path", "path_to_joint_solutions",
           "sampling_based_solve", "shortest_path", "shortest_path_csr", "solve", "solver_step"]
