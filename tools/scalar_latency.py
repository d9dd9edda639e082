import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
import acrobotics_b200 as ab
from acrobotics_b200.synthetic import readme_scene, readme_path, scene16
r = ab.Kuka(); sc = readme_scene(); s16 = scene16()
q = readme_path()[3]
for f, name in ((lambda: r.is_in_collision(q, sc), "is_in_collision(readme)"), (lambda: r.is_in_collision(q, s16), "is_in_collision(scene16)"), (lambda: r.fk(q), "fk"), (lambda: r.ik(r.fk(q)), "fk+ik"), (lambda: r.is_path_in_collision_discrete(q, q+0.5, sc), "path_discrete")):
    for _ in range(20): f()
    t0=time.perf_counter()
    for _ in range(500): f()
    print(name, "%.1f us/call" % ((time.perf_counter()-t0)/500*1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300): r.is_in_collision(q, sc)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
