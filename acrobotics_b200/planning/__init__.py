"""Sampling-based planner pieces that sit on the GPU hot path (reference
``src/acrobotics/planning/sampling_based.py``): joint solutions per path point and the
shortest path through the resulting layered graph."""
from .sampling_based import (COST_FUNCTIONS, JointPath, find_shortest_joint_path,
                             path_to_joint_solutions, shortest_path, shortest_path_csr)

__all__ = ["COST_FUNCTIONS", "JointPath", "find_shortest_joint_path", "path_to_joint_solutions",
           "shortest_path", "shortest_path_csr"]
