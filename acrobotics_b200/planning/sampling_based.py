"""Graph search of the sampling-based planner on the GPU.

Reference: ``find_shortest_joint_path`` (``src/acrobotics/planning/sampling_based.py:109-142``)
calls ``acrolib.dynamic_programming.shortest_path`` / ``shortest_path_with_state_cost`` with one
of ``acrolib.cost_functions.{norm_l1, norm_l2, sum_squared, weighted_sum_squared}``
(``:28-44``).  acrolib is not vendored in the reference tree; the semantics are restated from
these call sites (every node of layer i connects to every node of layer i+1, the result is a
dict with ``success``, ``path``, ``length``) - parity with acrolib itself is unpinned.
"""
import time
from copy import deepcopy

import numpy as np
import torch

from .. import _native, device
from ..path import path_pt as _path_pt
from ..path.path_pt import TolEulerPt, TolQuatPt
from .types import CostFuntionType, JointPath, Solution, SolveMethod

# cost(x, y) between joint vectors of consecutive layers
COST_FUNCTIONS = {"norm_l1": 0, "norm_l2": 1, "sum_squared": 2, "weighted_sum_squared": 3}


def shortest_path_csr(q, offsets, cost="sum_squared", weights=None, state_cost=None):
    """Cheapest chain through the layers of a CSR graph.

    ``q`` (M, ndof) rows (numpy or CUDA tensor), ``offsets`` (L+1,) row ranges of the layers.
    Returns ``(rows, length)``: ``rows[l]`` = index into ``q`` of the node chosen in layer l.
    Raises ``ValueError`` when a layer is empty (no path exists)."""
    device.require_cuda()
    host = not (isinstance(q, torch.Tensor) and q.is_cuda)
    qd = torch.as_tensor(q, dtype=torch.float64).to("cuda").contiguous()
    if qd.dim() != 2:
        raise ValueError("q must be (M, ndof)")
    off = np.ascontiguousarray(
        offsets.cpu().numpy() if isinstance(offsets, torch.Tensor) else offsets, dtype=np.int64)
    L = len(off) - 1
    if L < 1:
        raise ValueError("need at least one layer")
    if (np.diff(off) <= 0).any():
        raise ValueError("a layer has no nodes: no path exists")
    if off[0] != 0 or off[-1] != qd.shape[0]:
        raise ValueError("offsets do not cover q")
    code = COST_FUNCTIONS[cost] if isinstance(cost, str) else int(cost)
    ndof = int(qd.shape[1])
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (ndof,):
            raise ValueError("weights must have one entry per joint")
    sc = None
    if state_cost is not None:
        sc = torch.as_tensor(state_cost, dtype=torch.float64).to("cuda").contiguous()
        if sc.numel() != qd.shape[0]:
            raise ValueError("state_cost must have one entry per node")
    M = int(qd.shape[0])
    value = torch.empty(M, dtype=torch.float64, device="cuda")
    pred = torch.empty(M, dtype=torch.int32, device="cuda")
    off_dev = torch.empty(L + 1, dtype=torch.int64, device="cuda")
    rows = torch.empty(L, dtype=torch.int64, device="cuda")
    length = torch.empty(1, dtype=torch.float64, device="cuda")
    _native.check(
        _native.load().acro_shortest_path_f64(
            qd.data_ptr(), ndof, off.ctypes.data, L, code,
            w.ctypes.data if w is not None else None,
            sc.data_ptr() if sc is not None else None,
            value.data_ptr(), pred.data_ptr(), off_dev.data_ptr(), rows.data_ptr(),
            length.data_ptr(), device.current_stream_ptr(),
        ),
        "acro_shortest_path_f64",
    )
    torch.cuda.synchronize()  # `off` (host memory) was copied asynchronously
    return device.to_host(rows, host), float(length.item())


def shortest_path(joint_solutions, cost="sum_squared", weights=None, state_cost=None):
    """``acrolib.dynamic_programming.shortest_path[_with_state_cost]`` for a list of (n_i, ndof)
    arrays: ``{"success", "path", "length"}``; ``state_cost``: list of (n_i,) arrays."""
    if len(joint_solutions) == 0 or any(len(js) == 0 for js in joint_solutions):
        return {"success": False}
    layers = [np.asarray(js, dtype=np.float64).reshape(len(js), -1) for js in joint_solutions]
    q = np.concatenate(layers)
    offsets = np.concatenate([[0], np.cumsum([len(l) for l in layers])])
    sc = None if state_cost is None else np.concatenate([np.asarray(s, float) for s in state_cost])
    rows, length = shortest_path_csr(q, offsets, cost, weights, sc)
    return {"success": True, "path": [q[r] for r in rows], "length": length}


def find_shortest_joint_path(joint_solutions, cost="sum_squared", weights=None, state_cost=None):
    """``find_shortest_joint_path`` (planning/sampling_based.py:109-142).  Also accepts the
    reference's call shape ``(joint_solutions, setup, sampling_settings)``."""
    if hasattr(cost, "robot") and hasattr(weights, "search_strategy"):
        setup, s = cost, weights
        state_cost = None
        if s.use_state_cost:
            state_cost = calc_state_cost(joint_solutions, setup.robot, setup.path, s.state_cost_weight)
        cost = s.cost_function if s.cost_function is not None else "sum_squared"
        # only the weighted cost takes weights (the reference would hand them to any cost
        # function and fail inside acrolib for the two-argument ones)
        weights = s.weights if cost == "weighted_sum_squared" else None
    res = shortest_path(joint_solutions, cost, weights, state_cost)
    if not res["success"]:
        raise ValueError("Failed to find a shortest path in joint_solutions.")
    return JointPath(res["path"], res["length"])


def path_to_joint_solutions(current_path, setup_or_robot, settings_or_scene=None, settings=None):
    """Reference call shape ``(current_path, setup, sampling_settings)``
    (planning/sampling_based.py:94-106) or ``(path, robot, scene[, settings])``."""
    if hasattr(setup_or_robot, "robot"):
        return _path_pt.path_to_joint_solutions(current_path, setup_or_robot.robot,
                                                setup_or_robot.scene, settings_or_scene)
    return _path_pt.path_to_joint_solutions(current_path, setup_or_robot, settings_or_scene, settings)


_COST_OF_TYPE = {
    CostFuntionType.l1_norm: "norm_l1",
    CostFuntionType.l2_norm: "norm_l2",
    CostFuntionType.sum_squared: "sum_squared",
    CostFuntionType.weighted_sum_squared: "weighted_sum_squared",
}


def sampling_based_solve(planning_setup, settings):
    """The planner's outer loop (planning/sampling_based.py:22-79): joint solutions per path
    point, shortest path through them, then - for ``iterations > 1`` - tolerances shrunk around
    the solution found and the whole thing repeated."""
    assert settings.solve_method == SolveMethod.sampling_based
    s = settings.sampling_settings
    if s.cost_function is None:
        s.cost_function = _COST_OF_TYPE[settings.cost_function_type]
        if s.cost_function == "weighted_sum_squared":
            if s.weights is None:
                raise Exception("No weights specified in SamplingSettings for the weighted cost function.")
            s.weights = np.array(s.weights, dtype="float64")
    if callable(s.cost_function):
        raise NotImplementedError(
            "a Python cost function cannot run inside the shortest-path kernel: name one of "
            + ", ".join(COST_FUNCTIONS))

    current_path = [deepcopy(pt) for pt in planning_setup.path]
    current_joint_path = JointPath(None, None)
    all_paths, all_joint_paths = [], []
    start_time = time.time()
    for step in range(s.iterations):
        all_paths.append(deepcopy(current_path))
        current_path, current_joint_path = solver_step(step, current_path, current_joint_path,
                                                       planning_setup, s)
        all_joint_paths.append(deepcopy(current_joint_path))
    stop_time = time.time()
    extra_info = {"all_joint_paths": all_joint_paths, "all_paths": all_paths}
    return Solution(True, current_joint_path.joint_positions, current_joint_path.cost,
                    stop_time - start_time, extra_info)


def solver_step(step_count, prev_path, prev_sol, setup, sampling_settings):
    if step_count == 0:
        current_path = prev_path
    else:
        factor = getattr(sampling_settings, "tolerance_reduction_factor", 2.0)
        current_path = create_reduced_tolerance_path(prev_path, prev_sol.joint_positions,
                                                     setup.robot, factor)
    Q = path_to_joint_solutions(current_path, setup, sampling_settings)
    q_path = find_shortest_joint_path(Q, setup, sampling_settings)
    return current_path, q_path


def create_reduced_tolerance_path(current_path, current_joint_positions, robot,
                                  tolerance_reduction_factor=2.0):
    """New path points centred on the poses the current joint path reaches, tolerances shrunk
    but still containing the old reference (planning/sampling_based.py:145-152).  One batched
    FK for the whole path.  (The reference ignores the settings' reduction factor and always
    uses 2.0, ``create_new_pt``'s default; 2.0 is the default here too.)"""
    T = robot.fk_batch(np.asarray(current_joint_positions, dtype=np.float64))
    return [create_new_pt(T[i], q, pt, tolerance_reduction_factor)
            for i, (q, pt) in enumerate(zip(current_joint_positions, current_path))]


def create_new_pt(fk_transform, q, prev_pt, tolerance_reduction_factor=2.0):
    """planning/sampling_based.py:155-183"""
    if not isinstance(prev_pt, (TolQuatPt, TolEulerPt)):
        raise NotImplementedError()
    fk_transform = np.asarray(fk_transform, dtype=np.float64)
    new_tol = deepcopy(prev_pt.tol)
    ref_values = prev_pt.transform_to_rel_tolerance_deviation(fk_transform)
    for tolerance, ref in zip(new_tol, ref_values):
        tolerance.reduce_tolerance(tolerance_reduction_factor, ref)
    if isinstance(prev_pt, TolQuatPt):
        return TolQuatPt(fk_transform[:3, 3], fk_transform[:3, :3], new_tol[:3], new_tol[3])
    return TolEulerPt(fk_transform[:3, 3], fk_transform[:3, :3], new_tol[:3], new_tol[3:])


def calc_state_cost(joint_solutions, robot, initial_path, weight):
    """Welding cost ``weight * (r_x^2 + r_y^2)`` of every joint solution relative to its path
    point (planning/sampling_based.py:186-206); one batched FK per path point."""
    if not isinstance(initial_path[0], TolEulerPt):
        raise Exception("Current state cost only implemented for euler points.")
    state_cost = []
    for pt, qs in zip(initial_path, joint_solutions):
        T = robot.fk_batch(np.asarray(qs, dtype=np.float64))
        dev = pt.transform_to_rel_tolerance_deviation(T)
        cost = dev[:, 3] ** 2 + dev[:, 4] ** 2
        state_cost.append(weight * cost)
    return state_cost
