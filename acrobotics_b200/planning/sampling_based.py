"""Graph search of the sampling-based planner on the GPU.

Reference: ``find_shortest_joint_path`` (``src/acrobotics/planning/sampling_based.py:109-142``)
calls ``acrolib.dynamic_programming.shortest_path`` / ``shortest_path_with_state_cost`` with one
of ``acrolib.cost_functions.{norm_l1, norm_l2, sum_squared, weighted_sum_squared}``
(``:28-44``).  acrolib is not vendored in the reference tree; the semantics are restated from
these call sites (every node of layer i connects to every node of layer i+1, the result is a
dict with ``success``, ``path``, ``length``) - parity with acrolib itself is unpinned.
"""
import ctypes
from collections import namedtuple

import numpy as np
import torch

from .. import _native, device
from ..path.path_pt import path_to_joint_solutions  # noqa: F401  (reference keeps it here)

JointPath = namedtuple("JointPath", ["joint_positions", "cost"])

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
    """``find_shortest_joint_path`` (planning/sampling_based.py:109-142)."""
    res = shortest_path(joint_solutions, cost, weights, state_cost)
    if not res["success"]:
        raise ValueError("Failed to find a shortest path in joint_solutions.")
    return JointPath(res["path"], res["length"])
