"""Planner records (reference ``src/acrobotics/planning/types.py``)."""
import enum


class JointPath:
    def __init__(self, joint_positions, cost):
        self.joint_positions = joint_positions
        self.cost = cost

    def __iter__(self):  # unpacks like the namedtuple of round 1: (joint_positions, cost)
        return iter((self.joint_positions, self.cost))


class SolveMethod(enum.Enum):
    sampling_based = 0
    optimization_based = 1


class CostFuntionType(enum.Enum):  # (sic) the reference's spelling
    l1_norm = 0
    l2_norm = 1
    sum_squared = 3
    weighted_sum_squared = 4


class PlanningSetup:
    def __init__(self, robot, path, scene):
        self.robot = robot
        self.path = path
        self.scene = scene


class Solution:
    def __init__(self, success, joint_positions=None, path_cost=None, run_time=None, extra_info=None):
        self.success = success
        self.joint_positions = joint_positions
        self.path_cost = path_cost
        self.run_time = run_time
        self.extra_info = extra_info
