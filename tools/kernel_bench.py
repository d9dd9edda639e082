#!/usr/bin/env python
"""Per-kernel timing (CUDA events, device-resident inputs) of every kernel of the
hot path, one JSON line per kernel with its roofline view.

    python tools/kernel_bench.py [--n 16000000] [--reps 5] [--only k2,fk,...]

HBM-bound kernels report algorithmic GB/s against MEASURED_PEAKS.json's copy
bandwidth; the compute-bound fused FK+collision kernel reports configs/s and
W_full-equivalent TFLOP/s against the DFMA peak measured in the same run.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import acrobotics_b200 as ab  # noqa: E402
from acrobotics_b200 import _native  # noqa: E402
from acrobotics_b200.synthetic import random_configs_device, readme_scene, scene16  # noqa: E402


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, tot = 1e30, 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return best, tot / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--big", action="store_true", help="add the 64 M-configuration FK point")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    want = lambda name: not only or name in only  # noqa: E731

    lib = _native.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    n = args.n
    robot, scene = ab.Kuka(), scene16()
    q = random_configs_device(n, 6, seed=7)
    stream = torch.cuda.current_stream().cuda_stream
    out = []

    def emit(name, ms_best, ms_avg, units, bytes_per_unit=None, extra=None):
        rec = {"kernel": name, "n": units, "ms_best": round(ms_best, 4), "ms_avg": round(ms_avg, 4),
               "units_per_s": units / (ms_best * 1e-3)}
        if bytes_per_unit is not None:
            gbs = bytes_per_unit * units / (ms_best * 1e-3) / 1e9
            rec["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                               "frac": gbs / hbm_peak, "bytes_per_unit": bytes_per_unit}
        if extra:
            rec.update(extra)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    if want("peak"):
        for fp64 in (1, 0):
            tf, ms = _native.C.c_double(), _native.C.c_double()
            _native.check(lib.acro_measure_fma_peak(fp64, 4000, _native.C.byref(tf), _native.C.byref(ms)))
            print(json.dumps({"kernel": "fma_chain_f64" if fp64 else "fma_chain_f32",
                              "tflops": tf.value, "ms": ms.value}), flush=True)
            if fp64:
                peak64 = tf.value
    else:
        peak64 = None

    hk = robot._handle(kinematics_only=True)
    hfull = robot._handle()
    hs = scene.native_handle()

    if want("fk"):
        T = torch.empty((n, 4, 4), dtype=torch.float64, device="cuda")
        f = lambda: _native.check(lib.acro_fk_f64(hk, q.data_ptr(), n, 0, T.data_ptr(), 0, stream))  # noqa: E731
        emit("fk_f64", *timed(f, args.reps), n, 48 + 128)
        n1 = 1_000_000
        f = lambda: _native.check(lib.acro_fk_f64(hk, q.data_ptr(), n1, 0, T.data_ptr(), 0, stream))  # noqa: E731
        emit("fk_f64_1M", *timed(f, args.reps), n1, 48 + 128)
        if args.big:
            nb = 64_000_000  # steady-state point: 11 GB of traffic per launch
            qb = random_configs_device(nb, 6, seed=8)
            Tb = torch.empty((nb, 4, 4), dtype=torch.float64, device="cuda")
            f = lambda: _native.check(lib.acro_fk_f64(hk, qb.data_ptr(), nb, 0, Tb.data_ptr(), 0, stream))  # noqa: E731
            emit("fk_f64_64M", *timed(f, args.reps), nb, 48 + 128)
            del qb, Tb
        qs = q.t().contiguous()
        f = lambda: _native.check(lib.acro_fk_f64(hk, qs.data_ptr(), n, 1, T.data_ptr(), 0, stream))  # noqa: E731
        emit("fk_f64_soa", *timed(f, args.reps), n, 48 + 128)
        del qs
        q32 = q.float()
        T32 = torch.empty((n, 4, 4), dtype=torch.float32, device="cuda")
        f = lambda: _native.check(lib.acro_fk_f32(hk, q32.data_ptr(), n, 0, T32.data_ptr(), 0, stream))  # noqa: E731
        emit("fk_f32", *timed(f, args.reps), n, 24 + 64)
        del T32, T
        m = n // 4
        TA = torch.empty((m, 6, 4, 4), dtype=torch.float64, device="cuda")
        f = lambda: _native.check(lib.acro_fk_f64(hk, q.data_ptr(), m, 0, TA.data_ptr(), 1, stream))  # noqa: E731
        emit("fk_all_links_f64", *timed(f, args.reps), m, 48 + 768)
        del TA
        R = torch.empty((n, 6), dtype=torch.float64, device="cuda")
        f = lambda: _native.check(lib.acro_fk_rpy_f64(hk, q.data_ptr(), n, 0, R.data_ptr(), stream))  # noqa: E731
        emit("fk_rpy_f64", *timed(f, args.reps), n, 96)
        del R

    if want("jac"):
        hj = robot._handle(kinematics_only=True, as_constructed=True)
        J = torch.empty((n, 6, 6), dtype=torch.float64, device="cuda")
        f = lambda: _native.check(lib.acro_jac_pos_f64(hj, q.data_ptr(), n, 0, J.data_ptr(), stream))  # noqa: E731
        emit("jac_pos_f64", *timed(f, args.reps), n, 48 + 144)
        f = lambda: _native.check(lib.acro_jac_rpy_f64(hj, q.data_ptr(), n, 0, J.data_ptr(), stream))  # noqa: E731
        emit("jac_rpy_f64", *timed(f, args.reps), n, 48 + 288)
        del J

    if want("ik"):
        m = n // 4
        T = torch.empty((m, 4, 4), dtype=torch.float64, device="cuda")
        _native.check(lib.acro_fk_f64(hk, q.data_ptr(), m, 0, T.data_ptr(), 0, stream))
        sol = torch.empty((m, 8, 6), dtype=torch.float64, device="cuda")
        valid = torch.empty(m, dtype=torch.uint8, device="cuda")
        free = torch.empty(m, dtype=torch.uint8, device="cuda")
        f = lambda: _native.check(lib.acro_ik_kuka_f64(hk, T.data_ptr(), m, sol.data_ptr(), valid.data_ptr(), stream))  # noqa: E731
        emit("ik_kuka_f64", *timed(f, args.reps), m, 128 + 384 + 1)
        f = lambda: _native.check(lib.acro_ik_kuka_cc_f64(hfull, hs, T.data_ptr(), m, sol.data_ptr(), valid.data_ptr(), free.data_ptr(), stream))  # noqa: E731
        b, a = timed(f, args.reps)
        nsol = int(sum(bin(x).count("1") for x in valid[:100000].cpu().numpy().tolist()) / 100000 * m)
        emit("ik_kuka_cc_f64", b, a, m, None, {"ik_solutions_checked_per_s": nsol / (b * 1e-3)})
        del T, sol

    if want("k2"):
        words = (n + 31) // 32
        mask = torch.zeros(words, dtype=torch.int32, device="cuda")
        near = torch.zeros(words, dtype=torch.int32, device="cuda")

        def k2_extra(ms):
            cps = n / (ms * 1e-3)
            tfl = 26190 * cps / 1e12
            e = {"configs_per_s": cps, "w_full_tflops": tfl}
            if peak64:
                e["frac_of_fp64_peak_wfull"] = tfl / peak64
            return e

        readme = readme_scene()  # keep the Scene alive: it owns the native handle
        for label, sc_h in (("scene16", hs), ("readme2", readme.native_handle()), ("self_only", None)):
            f = lambda: _native.check(lib.acro_fk_cc_f64(hfull, sc_h, q.data_ptr(), n, 0, mask.data_ptr(), None, 0.0, stream))  # noqa: E731
            b, a = timed(f, args.reps)
            emit(f"fk_cc_f64[{label}]", b, a, n, 48.125, k2_extra(b))
        f = lambda: _native.check(lib.acro_fk_cc_f64(hfull, hs, q.data_ptr(), n, 0, mask.data_ptr(), near.data_ptr(), 1e-6, stream))  # noqa: E731
        b, a = timed(f, args.reps)
        emit("fk_cc_f64_margin[scene16]", b, a, n, 48.25, k2_extra(b))
        qs = q.t().contiguous()
        f = lambda: _native.check(lib.acro_fk_cc_f64(hfull, hs, qs.data_ptr(), n, 1, mask.data_ptr(), None, 0.0, stream))  # noqa: E731
        b, a = timed(f, args.reps)
        emit("fk_cc_f64_soa[scene16]", b, a, n, 48.125, k2_extra(b))
        del qs
        q32 = q.float()
        f = lambda: _native.check(lib.acro_fk_cc_f32(hfull, hs, q32.data_ptr(), n, 0, mask.data_ptr(), stream))  # noqa: E731
        b, a = timed(f, args.reps)
        emit("fk_cc_f32[scene16]", b, a, n, 24.125, k2_extra(b))

    if want("planner"):
        # BASELINE config 5: 10k-point tolerance path x 400 pose samples x <= 8 IK branches
        P, S = 10_000, 400
        m = P * S
        T = torch.empty((m, 4, 4), dtype=torch.float64, device="cuda")
        _native.check(lib.acro_fk_f64(hk, q.data_ptr(), min(m, n), 0, T.data_ptr(), 0, stream))
        if m > n:
            T[n:] = T[: m - n]
        q_all = torch.empty((m, 8, 6), dtype=torch.float64, device="cuda")
        valid = torch.empty(m, dtype=torch.uint8, device="cuda")
        free = torch.empty(m, dtype=torch.uint8, device="cuda")
        q_free = torch.empty((8 * m, 6), dtype=torch.float64, device="cuda")
        offsets = torch.zeros(P + 1, dtype=torch.int64, device="cuda")
        f = lambda: _native.check(lib.acro_joint_solutions_csr_f64(  # noqa: E731
            hfull, hs, T.data_ptr(), P, S, q_all.data_ptr(), valid.data_ptr(), free.data_ptr(),
            q_free.data_ptr(), 8 * m, offsets.data_ptr(), stream))
        b, a = timed(f, args.reps)
        total = int(offsets[-1].item())
        emit("joint_solutions_csr_f64[10k x 400]", b, a, m, None,
             {"poses_per_s": m / (b * 1e-3), "ik_solutions_checked_per_s": 8 * m / (b * 1e-3),
              "collision_free_solutions": total})
        # shortest path through the resulting layered graph (10k layers)
        off_host = offsets.cpu().numpy().astype(np.int64)
        if (np.diff(off_host) > 0).all():
            M = int(off_host[-1])
            value = torch.empty(M, dtype=torch.float64, device="cuda")
            pred = torch.empty(M, dtype=torch.int32, device="cuda")
            off_dev = torch.empty(P + 1, dtype=torch.int64, device="cuda")
            rows = torch.empty(P, dtype=torch.int64, device="cuda")
            length = torch.empty(1, dtype=torch.float64, device="cuda")
            f = lambda: _native.check(lib.acro_shortest_path_f64(  # noqa: E731
                q_free.data_ptr(), 6, off_host.ctypes.data, P, 2, None, None, value.data_ptr(),
                pred.data_ptr(), off_dev.data_ptr(), rows.data_ptr(), length.data_ptr(), stream))
            b, a = timed(f, max(2, args.reps // 2), warm=1)
            edges = float((np.diff(off_host)[:-1].astype(np.float64) * np.diff(off_host)[1:]).sum())
            emit("shortest_path_f64[10k layers]", b, a, M, None,
                 {"edges": edges, "edges_per_s": edges / (b * 1e-3),
                  "fp64_tflops": edges * 18 / (b * 1e-3) / 1e12, "length": float(length.item())})
        del T, q_all, q_free

    if want("path"):
        e = n // 16
        qs = q[:e].contiguous()
        qg = (qs + 0.3 * torch.randn_like(qs)).contiguous()
        mask = torch.zeros((e + 31) // 32, dtype=torch.int32, device="cuda")
        f = lambda: _native.check(lib.acro_path_cc_f64(hfull, hs, qs.data_ptr(), qg.data_ptr(), e, 0.1, mask.data_ptr(), stream))  # noqa: E731
        b, a = timed(f, args.reps)
        steps = torch.ceil(torch.linalg.norm(qg - qs, dim=1) / 0.1).sum().item()
        emit("path_cc_f64[scene16]", b, a, e, None, {"interpolants_upper_bound_per_s": steps / (b * 1e-3)})
        # K7: swept collision of the same segments (ten FCL motion samples per box, all FP64)
        f = lambda: _native.check(lib.acro_swept_cc_f64(hfull, hs, qs.data_ptr(), qg.data_ptr(), e, 0,  # noqa: E731
                                                        mask.data_ptr(), None, 0.0, stream))
        b, a = timed(f, args.reps)
        emit("swept_cc_f64[scene16]", b, a, e, None, {"pose_checks_per_s": 10 * e / (b * 1e-3)})


if __name__ == "__main__":
    main()
