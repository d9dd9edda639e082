#!/usr/bin/env python
"""K7 (swept collision) timing: `python tools/swept_bench.py [--n 4000000] [--profile]`.
CUDA events on the launching stream, inputs resident in HBM and larger than L2 (2 x 48 B per
pair).  Prints one JSON line per workload; --profile runs each kernel twice only (for ncu)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200.synthetic import random_configs_device, scene16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4_000_000)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
n = args.n
scene = scene16()
g = np.load(os.path.join(ROOT, "tests", "golden", "ccd.npz"))
kuka = ab.Kuka()
torch_kuka = ab.Kuka()
torch_kuka.tool = ab.Tool([ab.Box(*s) for s in g["tool_sides"]], list(g["tool_tfs"]), g["tool_tip"])
qs = random_configs_device(n, 6, seed=7)
for name, robot, span in (("kuka", kuka, 0.2), ("kuka", kuka, 1.0), ("kuka+torch2", torch_kuka, 1.0)):
    qg = qs + (torch.rand_like(qs) * 2 - 1) * span
    f = lambda: robot.is_path_in_collision_batch(qs, qg, scene, packed=True)  # noqa: E731
    for _ in range(1 if args.profile else 3):
        out = f()
    torch.cuda.synchronize()
    reps = 1 if args.profile else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bits = np.unpackbits(out.cpu().numpy().view(np.uint8), bitorder="little")[:n]
    print(json.dumps({"kernel": "swept_cc_kernel", "robot": name, "scene_boxes": 16, "pairs": n,
                      "dq_span": span, "samples": 10, "ms": round(ms, 4),
                      "pairs_per_s": n / ms * 1e3, "pose_checks_per_s": 10 * n / ms * 1e3,
                      "hit_rate": float(bits.mean())}))
