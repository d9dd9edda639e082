"""Solver settings (reference ``src/acrobotics/planning/settings.py``).  Only the
sampling-based solver is on the GPU path; ``OptSettings`` is kept as a record so that user
code constructs, but the optimisation back-end (casadi / IPOPT) is out of scope."""
from ..path.sampling import SamplingSetting
from .types import CostFuntionType, SolveMethod


class OptSettings:
    def __init__(self, q_init=None, max_iters=None, weights=None, con_objective_weight=0.0):
        self.q_init = q_init
        self.weights = weights
        self.con_objective_weight = con_objective_weight
        self.max_iters = 100 if max_iters is None else max_iters


class SolverSettings:
    def __init__(self, solve_method: SolveMethod, cost_function_type: CostFuntionType,
                 sampling_settings: SamplingSetting = None, opt_settings: OptSettings = None):
        self.solve_method = solve_method
        self.cost_function_type = cost_function_type
        if solve_method == SolveMethod.sampling_based:
            assert sampling_settings is not None
            self.sampling_settings = sampling_settings
        elif solve_method == SolveMethod.optimization_based:
            assert opt_settings is not None
            self.opt_settings = opt_settings
