#!/usr/bin/env python
"""A/B timing of the production K2 kernel (bench workload: Kuka x scene16, random configs).

    ACRO_B200_LIB=<variant .so> python tools/k2_quick.py [--n 100000000] [--reps 5]

Prints one JSON line: best / average ms, configs/s and a checksum of the verdict mask (equal
checksums across variants = identical verdicts on the same inputs)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200 import _native  # noqa: E402
from acrobotics_b200.synthetic import random_configs_device, scene16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--seed", type=int, default=1000)
ap.add_argument("--tag", default=os.path.basename(os.environ.get("ACRO_B200_LIB", "default")))
args = ap.parse_args()
lib = _native.load()
robot, scene = ab.Kuka(), scene16()
n = args.n
q = random_configs_device(n, 6, seed=args.seed)
mask = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rh, sh = robot._handle(), scene.native_handle()


def f():
    _native.check(lib.acro_fk_cc_f64(rh, sh, q.data_ptr(), n, 0, mask.data_ptr(), None, 0.0, st))


for _ in range(3):
    f()
torch.cuda.synchronize()
ms = []
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f()
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
chk = int(mask.to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF
print(json.dumps({"tag": args.tag, "n": n, "ms_best": min(ms), "ms_avg": sum(ms) / len(ms),
                  "configs_per_s": n / (min(ms) * 1e-3), "mask_checksum": chk}), flush=True)
