#!/usr/bin/env python
"""Small run of every kernel (for compute-sanitizer memcheck / racecheck on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200 import _native  # noqa: E402
from acrobotics_b200.path import joint_solutions_csr  # noqa: E402
from acrobotics_b200.planning import shortest_path  # noqa: E402
from acrobotics_b200.synthetic import random_configs, readme_scene, scene16  # noqa: E402

k = ab.Kuka()
rail = ab.KukaOnRail()
n = 9001
q = random_configs(n, 6, seed=1)
for scene in (scene16(), readme_scene(), scene16(seed=2, n_boxes=40), ab.Scene([], [])):
    for variant in ("filter", "wave", "simple"):
        _native.set_option("k2_variant", variant)
        k.is_in_collision_batch(q, scene)
    _native.set_option("k2_variant", "filter")
    k.is_in_collision_batch(q, scene, return_near=True)
    k.is_in_collision_batch(q[:77], scene)
rail.is_in_collision_batch(random_configs(n, 7, seed=2), scene16())
k.is_in_self_collision_batch(q)
T = k.fk_batch(q)
k.fk_batch(q, all_links=True)
k.fk_batch(q, dtype="f32")
k.fk_rpy_batch(q)
k.jacobian_position_batch(q)
k.jacobian_rpy_batch(q)
k.ik_batch(T)
k.ik_filtered_batch(T, scene16())
k.ik_filtered_batch(T[:100], scene16())
qg = q + 0.3 * np.random.default_rng(0).normal(size=q.shape)
k.is_path_in_collision_discrete_batch(q, qg, scene16(), 0.1)
k.is_path_in_collision_discrete_batch(q[:50], qg[:50], scene16(), 0.1)
qf, off = joint_solutions_csr(k, T[:9000].reshape(90, 100, 4, 4), scene16())
layers = [qf[off[i]:off[i + 1]] for i in range(90) if off[i + 1] > off[i]]
print(shortest_path(layers, "sum_squared")["length"], len(qf))
print("sanitize smoke done")
