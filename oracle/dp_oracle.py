"""CPU restatement of the layered-graph shortest path (TEST INFRASTRUCTURE).

Follows the call sites of ``acrolib.dynamic_programming.shortest_path`` /
``shortest_path_with_state_cost`` in reference ``src/acrobotics/planning/sampling_based.py:109-142``
and the four cost functions named at ``:28-44``.  acrolib itself is un-vendored: parity unpinned.
"""
import numpy as np


def edge_costs(a, b, cost, weights=None):
    """(na, nb) matrix of cost(a_i, b_j)."""
    d = a[:, None, :] - b[None, :, :]
    if cost == "norm_l1":
        return np.abs(d).sum(-1)
    if cost == "norm_l2":
        return np.sqrt((d * d).sum(-1))
    if cost == "sum_squared":
        return (d * d).sum(-1)
    if cost == "weighted_sum_squared":
        return (np.asarray(weights) * d * d).sum(-1)
    raise ValueError(cost)


def shortest_path(layers, cost="sum_squared", weights=None, state_cost=None):
    if len(layers) == 0 or any(len(l) == 0 for l in layers):
        return {"success": False}
    V = np.zeros(len(layers[0])) if state_cost is None else np.asarray(state_cost[0], float).copy()
    preds = []
    for i in range(1, len(layers)):
        C = V[:, None] + edge_costs(layers[i - 1], layers[i], cost, weights)
        p = C.argmin(axis=0)  # first minimum = lowest index
        V = C[p, np.arange(C.shape[1])]
        if state_cost is not None:
            V = V + np.asarray(state_cost[i], float)
        preds.append(p)
    k = int(V.argmin())
    length = float(V[k])
    idx = [k]
    for p in reversed(preds):
        k = int(p[k])
        idx.append(k)
    idx.reverse()
    return {"success": True, "path": [layers[i][j] for i, j in enumerate(idx)], "length": length,
            "index": idx}
