#!/usr/bin/env python
"""Launch every kernel of the path besides K2 a few times (for an ncu capture that stays small):

    ncu --set full --clock-control none -k regex:<pattern> -s <skip> -c <count> -o out \
        python tools/profile_streaming.py [--n 4000000] [--only fk,fk32,jac,ik,path,planner,dp]

Each kernel runs twice: the first launch is the warm-up, profile the second."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200 import _native  # noqa: E402
from acrobotics_b200.synthetic import random_configs_device, scene16  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4_000_000)
ap.add_argument("--only", default="fk,fk32,fkall,rpy,jac,ik,path,planner,dp")
args = ap.parse_args()
want = set(args.only.split(","))
lib = _native.load()
n = args.n
robot, scene = ab.Kuka(), scene16()
hk, hf, hs = robot._handle(kinematics_only=True), robot._handle(), scene.native_handle()
hj = robot._handle(kinematics_only=True, as_constructed=True)
q = random_configs_device(n, 6, seed=7)
st = torch.cuda.current_stream().cuda_stream
ck = _native.check


def twice(f):
    f()
    torch.cuda.synchronize()
    f()
    torch.cuda.synchronize()


if "fk" in want:
    T = torch.empty((n, 4, 4), dtype=torch.float64, device="cuda")
    twice(lambda: ck(lib.acro_fk_f64(hk, q.data_ptr(), n, 0, T.data_ptr(), 0, st)))
    twice(lambda: ck(lib.acro_fk_f64(hk, q.data_ptr(), 1_000_000, 0, T.data_ptr(), 0, st)))
    del T
if "fk32" in want:
    q32 = q.float()
    T32 = torch.empty((n, 4, 4), dtype=torch.float32, device="cuda")
    twice(lambda: ck(lib.acro_fk_f32(hk, q32.data_ptr(), n, 0, T32.data_ptr(), 0, st)))
    del T32, q32
if "fkall" in want:
    m = n // 4
    TA = torch.empty((m, 6, 4, 4), dtype=torch.float64, device="cuda")
    twice(lambda: ck(lib.acro_fk_f64(hk, q.data_ptr(), m, 0, TA.data_ptr(), 1, st)))
    del TA
if "rpy" in want:
    R = torch.empty((n, 6), dtype=torch.float64, device="cuda")
    twice(lambda: ck(lib.acro_fk_rpy_f64(hk, q.data_ptr(), n, 0, R.data_ptr(), st)))
    del R
if "jac" in want:
    J = torch.empty((n, 6, 6), dtype=torch.float64, device="cuda")
    twice(lambda: ck(lib.acro_jac_pos_f64(hj, q.data_ptr(), n, 0, J.data_ptr(), st)))
    twice(lambda: ck(lib.acro_jac_rpy_f64(hj, q.data_ptr(), n, 0, J.data_ptr(), st)))
    del J
if "ik" in want:
    m = n // 4
    T = torch.empty((m, 4, 4), dtype=torch.float64, device="cuda")
    ck(lib.acro_fk_f64(hk, q.data_ptr(), m, 0, T.data_ptr(), 0, st))
    sol = torch.empty((m, 8, 6), dtype=torch.float64, device="cuda")
    valid = torch.empty(m, dtype=torch.uint8, device="cuda")
    twice(lambda: ck(lib.acro_ik_kuka_f64(hk, T.data_ptr(), m, sol.data_ptr(), valid.data_ptr(), st)))
    del T, sol
if "path" in want:
    e = n // 8
    qs = q[:e].contiguous()
    qg = (qs + 0.3 * torch.randn_like(qs)).contiguous()
    mask = torch.zeros((e + 31) // 32, dtype=torch.int32, device="cuda")
    twice(lambda: ck(lib.acro_path_cc_f64(hf, hs, qs.data_ptr(), qg.data_ptr(), e, 0.1, mask.data_ptr(), st)))
if "planner" in want or "dp" in want:
    P, S = 2000, 400
    m = P * S
    T = torch.empty((m, 4, 4), dtype=torch.float64, device="cuda")
    ck(lib.acro_fk_f64(hk, q.data_ptr(), min(m, n), 0, T.data_ptr(), 0, st))
    q_all = torch.empty((m, 8, 6), dtype=torch.float64, device="cuda")
    valid = torch.empty(m, dtype=torch.uint8, device="cuda")
    free = torch.empty(m, dtype=torch.uint8, device="cuda")
    q_free = torch.empty((8 * m, 6), dtype=torch.float64, device="cuda")
    offsets = torch.zeros(P + 1, dtype=torch.int64, device="cuda")
    twice(lambda: ck(lib.acro_joint_solutions_csr_f64(hf, hs, T.data_ptr(), P, S, q_all.data_ptr(),
                                                      valid.data_ptr(), free.data_ptr(), q_free.data_ptr(),
                                                      8 * m, offsets.data_ptr(), st)))
    if "dp" in want:
        off_host = offsets.cpu().numpy().astype(np.int64)
        M = int(off_host[-1])
        value = torch.empty(M, dtype=torch.float64, device="cuda")
        pred = torch.empty(M, dtype=torch.int32, device="cuda")
        off_dev = torch.empty(P + 1, dtype=torch.int64, device="cuda")
        rows = torch.empty(P, dtype=torch.int64, device="cuda")
        length = torch.empty(1, dtype=torch.float64, device="cuda")
        ck(lib.acro_shortest_path_f64(q_free.data_ptr(), 6, off_host.ctypes.data, 40, 2, None, None,
                                      value.data_ptr(), pred.data_ptr(), off_dev.data_ptr(), rows.data_ptr(),
                                      length.data_ptr(), st))
        torch.cuda.synchronize()
print("done")
