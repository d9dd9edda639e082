"""The reference's end-to-end planner tests (tests/test_planning_sampling_based.py:30-297)
through this package: same robots, scenes, paths and settings, the same assertions on the
solution.  Orientations are rotation matrices here (the reference builds them with
pyquaternion from the same axis / angle)."""
import numpy as np
import pytest

import acrobotics_b200 as ab
from acrobotics_b200.path import SampleMethod
from acrobotics_b200.path.factory import axis_angle_matrix
from acrobotics_b200.planning import calc_state_cost, create_new_pt

pytestmark = pytest.mark.gpu


def _table(z):
    tf = np.array([[1, 0, 0, 0.80], [0, 1, 0, 0.00], [0, 0, 1, z], [0, 0, 0, 1.0]])
    return ab.Scene([ab.Box(0.5, 0.5, 0.1)], [tf])


def test_tol_quat_pt_with_weights(gpu):
    # reference :30-87: orientation-free points, INCREMENTAL strategy, two iterations, weights
    R = axis_angle_matrix([1, 0, 0], np.pi)
    path = [ab.TolQuatPt([0.8, s * 0.2 + (1 - s) * (-0.2), 0.2], R, [ab.NoTolerance()] * 3,
                         ab.QuaternionTolerance(2.0)) for s in np.linspace(0, 1, 3)]
    robot = ab.Kuka()
    setup = ab.PlanningSetup(robot, path, _table(0.12))
    settings = ab.SamplingSetting(ab.SearchStrategy.INCREMENTAL, sample_method=SampleMethod.random_uniform,
                                  num_samples=500, iterations=2, tolerance_reduction_factor=2,
                                  weights=[10.0, 5.0, 1.0, 1.0, 1.0, 1.0])
    solve_set = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.weighted_sum_squared,
                                  sampling_settings=settings)
    sol = ab.solve(setup, solve_set)
    assert sol.success and len(sol.extra_info["all_joint_paths"]) == 2
    for qi, s in zip(sol.joint_positions, np.linspace(0, 1, 3)):
        assert np.allclose(robot.fk(qi)[:3, 3], [0.8, s * 0.2 + (1 - s) * (-0.2), 0.2], atol=1e-7)
        assert not robot.is_in_collision(qi, setup.scene)


def test_tol_position_pt_planning_problem(gpu):
    # reference :95-139
    robot = ab.Kuka()
    quat = axis_angle_matrix([1, 0, 0], np.pi)
    tolerance = [ab.NoTolerance(), ab.SymmetricTolerance(0.05, 10), ab.NoTolerance()]
    first_point = ab.TolPositionPt(np.array([0.9, -0.2, 0.2]), quat, tolerance)
    path = ab.create_arc(first_point, np.array([0.9, 0.0, 0.2]), np.array([0, 0, 1]), 2 * np.pi, 5)
    settings = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.sum_squared,
                                 sampling_settings=ab.SamplingSetting(ab.SearchStrategy.GRID, iterations=1))
    sol = ab.solve(ab.PlanningSetup(robot, path, _table(0.12)), settings)
    assert sol.success
    for qi, pt in zip(sol.joint_positions, path):
        pos_error = pt.rotation_matrix.T @ (robot.fk(qi)[:3, 3] - pt.pos)
        assert abs(pos_error[0]) < 1e-7 and abs(pos_error[2]) < 1e-7
        assert -(0.05 + 1e-6) <= pos_error[1] <= 0.05 + 1e-6


def _euler_problem():
    robot = ab.Kuka()
    quat = axis_angle_matrix([1, 0, 0], -3 * np.pi / 4)
    rot_tol = [ab.NoTolerance(), ab.SymmetricTolerance(np.pi / 4, 20), ab.SymmetricTolerance(np.pi, 20)]
    first_point = ab.TolEulerPt(np.array([0.9, -0.1, 0.2]), quat, 3 * [ab.NoTolerance()], rot_tol)
    path = ab.create_arc(first_point, np.array([0.9, 0.0, 0.2]), np.array([0, 0, 1]), 2 * np.pi, 5)
    return robot, path, _table(0.0)


def test_euler_pt_planning_problem_and_state_cost(gpu):
    # reference :143-195 and :198-239
    robot, path, scene = _euler_problem()
    lengths = {}
    for use_state_cost in (False, True):
        s = ab.SamplingSetting(ab.SearchStrategy.GRID, iterations=1, tolerance_reduction_factor=2,
                               use_state_cost=use_state_cost, state_cost_weight=10.0)
        settings = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.sum_squared,
                                     sampling_settings=s)
        sol = ab.solve(ab.PlanningSetup(robot, path, scene), settings)
        assert sol.success and len(sol.joint_positions) == 5
        for qi, pt in zip(sol.joint_positions, path):
            assert np.allclose(robot.fk(qi)[:3, 3], pt.pos, atol=1e-7)
            assert not robot.is_in_collision(qi, scene)
            dev = pt.transform_to_rel_tolerance_deviation(robot.fk(qi))
            assert abs(dev[4]) <= np.pi / 4 + 1e-6 and abs(dev[3]) < 1e-6
        lengths[use_state_cost] = sol.path_cost
    assert lengths[True] >= lengths[False] - 1e-9  # the state cost can only add
    # the state cost itself: weight * (r_x^2 + r_y^2) of the solution's deviation
    js = [np.array([q]) for q in sol.joint_positions]
    sc = calc_state_cost(js, robot, path, 10.0)
    for c, q, pt in zip(sc, sol.joint_positions, path):
        d = pt.transform_to_rel_tolerance_deviation(robot.fk(q))
        assert np.allclose(c, 10.0 * (d[3] ** 2 + d[4] ** 2))


def test_two_iterations_shrink_the_tolerances(gpu):
    robot, path, scene = _euler_problem()
    s = ab.SamplingSetting(ab.SearchStrategy.GRID, iterations=2, tolerance_reduction_factor=2)
    settings = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.sum_squared,
                                 sampling_settings=s)
    sol = ab.solve(ab.PlanningSetup(robot, path, scene), settings)
    first, second = sol.extra_info["all_paths"]
    for a, b in zip(first, second):
        for ta, tb in zip(a.tol, b.tol):
            assert (tb.upper - tb.lower) <= (ta.upper - ta.lower) + 1e-12
    c1, c2 = (jp.cost for jp in sol.extra_info["all_joint_paths"])
    assert np.isfinite(c1) and np.isfinite(c2)
    # create_new_pt centres the new point on the pose reached (reference :155-183)
    q = sol.extra_info["all_joint_paths"][0].joint_positions[0]
    new = create_new_pt(robot.fk(q), q, first[0])
    assert np.allclose(new.pos, robot.fk(q)[:3, 3]) and isinstance(new, ab.TolEulerPt)
    with pytest.raises(NotImplementedError):
        create_new_pt(robot.fk(q), q, ab.TolPositionPt([0, 0, 0], np.eye(3), [ab.NoTolerance()] * 3))


def test_min_incremental_and_exceptions(gpu):
    # reference :242-297
    settings = ab.SamplingSetting(ab.SearchStrategy.INCREMENTAL, sample_method=SampleMethod.random_uniform,
                                  num_samples=500, iterations=2, tolerance_reduction_factor=2)
    solve_set = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.weighted_sum_squared,
                                  sampling_settings=settings)
    with pytest.raises(Exception) as e:
        ab.solve(ab.PlanningSetup(None, None, None), solve_set)
    assert str(e.value) == "No weights specified in SamplingSettings for the weighted cost function."
    robot = ab.Kuka()
    quat = axis_angle_matrix([1, 0, 0], -3 * np.pi / 4)
    path = [ab.TolPositionPt(np.array([1000, 0, 0]), quat, 3 * [ab.NoTolerance()])]
    settings = ab.SamplingSetting(ab.SearchStrategy.INCREMENTAL, sample_method=SampleMethod.random_uniform,
                                  num_samples=50)
    solve_set2 = ab.SolverSettings(ab.SolveMethod.sampling_based, ab.CostFuntionType.sum_squared,
                                   sampling_settings=settings)
    with pytest.raises(Exception) as e:
        ab.solve(ab.PlanningSetup(robot, path, ab.Scene([], [])), solve_set2)
    assert str(e.value) == "No valid joint solutions for path point 0."
    # MIN_INCREMENTAL on the euler problem: at least the desired number of solutions per point,
    # all of them collision free, and the search stops (no exception) well before the cap
    robot, path, scene = _euler_problem()
    s = ab.SamplingSetting(ab.SearchStrategy.MIN_INCREMENTAL, sample_method=SampleMethod.deterministic_uniform,
                           desired_num_samples=12, max_search_iters=400)
    from acrobotics_b200.path import path_to_joint_solutions

    Q = path_to_joint_solutions(path, robot, scene, s)
    for js in Q:
        assert len(js) >= 12 and not robot.is_in_collision_batch(js, scene).any()
    one = path[0].to_joint_solutions(robot, s, scene)
    assert len(one) >= 12
    s.max_search_iters = 1
    s.desired_num_samples = 10_000
    with pytest.raises(Exception, match="Maximum iterations reached"):
        path_to_joint_solutions(path, robot, scene, s)
    with pytest.raises(NotImplementedError):
        ab.solve(None, ab.SolverSettings(ab.SolveMethod.optimization_based, ab.CostFuntionType.sum_squared,
                                         opt_settings=ab.OptSettings()))


def test_generic_robot_goes_through_its_own_ik(gpu):
    """A 6R chain that is not a Kuka must never get the Kuka's closed form:
    the native entry point refuses it and the path layer falls back to robot.ik."""
    from acrobotics_b200 import _native
    from acrobotics_b200.catalog import Kuka

    class NotAKuka(Kuka):
        pass

    odd = NotAKuka(a1=0.18, a2=0.6, d4=0.62, d6=0.115)
    odd.links[2] = type(odd.links[2])(odd.links[2].dh._replace(a=0.1), odd.links[2].joint_type,
                                      odd.links[2].geometry)
    with pytest.raises(_native.NativeError, match="Kuka DH pattern"):
        odd.ik_batch(np.eye(4)[None])

    class Stub:  # generic robot: own ik, batched collision call of the product
        ndof = 6

        def __init__(self):
            self.k = ab.Kuka()

        def ik(self, T):
            return self.k.ik(T)

        def is_in_collision_batch(self, q, scene):
            return self.k.is_in_collision_batch(q, scene)

    robot, path, scene = _euler_problem()
    want = path[1].to_joint_solutions(robot, None, scene)
    got = path[1].to_joint_solutions(Stub(), None, scene)
    assert got.shape == want.shape and np.allclose(got, want, atol=1e-12)
