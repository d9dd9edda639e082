#!/usr/bin/env python
"""Benchmark of the hot path: fused FK -> collision check (K2) of the Kuka
against the 16-box synthetic scene (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--configs M] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of M configurations per
GPU (default 100 M, resident in HBM: 4.8 GB, far larger than the 126 MB L2, so
every step streams from DRAM).  N > 1 is launched by torchrun, one rank per
GPU; every rank checks its own shard and the bit masks are all-gathered (NCCL)
inside the timed step.  Rank 0 prints ONE JSON line.

``--impl reference`` times the CPU arm: the reference's per-configuration
Python path (oracle/scalar_port.py - the reference tree does not exist on the
GPU box and python-fcl is not installable) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collision-checked configs/sec (FK+CC)"
UNIT = "configs/s"
WORKLOAD = "Kuka 6-DOF fused FK+collision vs 16-box synthetic Scene (seed 0), q~U[-pi,pi)^6"
# algorithmic work per configuration, FMA = 2 flop, no culling:
# FK 63 flop/joint * 6 + 18 flop/link box * 6 = 486; 6*16 + 6 = 102 box pairs * 252 flop
W_FULL_FLOP = 486 + 102 * 252
BYTES_PER_CONFIG = 48.125  # 6 f64 in, 1 bit out
K2_KERNEL = "acro::fk_cc_filter_kernel<6,uint32_t> (+ acro::fk_cc_fixup_kernel<6>, ~2 % of the step)"


def _load_k2_profile():
    """Per-configuration counters of the production K2 kernel from its last ncu
    --set full capture (written by tools/ncu_k2_profile.py, tracked under profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "k2_profile.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


K2_PROFILE = _load_k2_profile()


# ----------------------------------------------------------------------------
# CPU arm (oracle port; allowed here as the timed baseline only)
# ----------------------------------------------------------------------------
def _cpu_setup():
    import acrobotics_b200 as ab
    from acrobotics_b200.synthetic import scene16
    from oracle import np_oracle as no

    spec = no.spec_from_robot(ab.Kuka())
    scene = no.scene_arrays(scene16())
    return spec, scene


def cpu_port_throughput(target_s, cores, seed=0):
    """Configs/s of the per-configuration port over `cores` processes on a sample
    sized for about `target_s` seconds.  Returns (value, n_sample, seconds)."""
    import multiprocessing as mp

    import numpy as np

    from oracle import scalar_port

    spec, scene = _cpu_setup()
    rng = np.random.default_rng(seed)
    pilot = rng.uniform(-np.pi, np.pi, (300, 6))
    t0 = time.perf_counter()
    scalar_port.run_chunk((spec, scene, pilot))
    rate1 = len(pilot) / (time.perf_counter() - t0)
    n = int(max(cores * 200, min(rate1 * cores * target_s, 2_000_000)))
    q = rng.uniform(-np.pi, np.pi, (n, 6))
    chunks = np.array_split(q, cores * 4)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(scalar_port.run_chunk, [(spec, scene, c[:8]) for c in chunks[:cores]])  # warm
        t0 = time.perf_counter()
        pool.map(scalar_port.run_chunk, [(spec, scene, c) for c in chunks])
        dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step_s = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    n_sample = 0
    for i in range(args.warmup + args.steps):
        v, n_sample, dt = cpu_port_throughput(per_step_s, cores, seed=i)
        if i >= args.warmup:
            vals.append((v, dt))
    value = statistics.mean(v for v, _ in vals)
    sample = (f"{n_sample} random configurations per step over {cores} processes "
              f"(oracle/scalar_port.py: reference per-configuration Python path, C box-box)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.mean(dt for _, dt in vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, row in self.rows:
            if not (t0 <= t <= t1 + 0.1):
                continue
            f = [x.strip() for x in row.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(smax) if smax else None,
            "power_w": statistics.median(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (rank 0, N=1 only), before any CUDA context exists
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, n_sample, dt = cpu_port_throughput(args.cpu_seconds, cores)
        cpu_baseline = {
            "value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"{n_sample} random configurations in {dt:.1f} s over {cores} processes; "
                       "oracle/scalar_port.py = the reference's per-configuration Python path "
                       "(numpy 4x4 chain, one native box-box call per pair)"),
        }

    if cpu_baseline is not None:
        # context: the vectorised numpy oracle on one core (SURVEY 8d, item ii)
        import numpy as _np

        from oracle import np_oracle as _no

        _spec, _scene = _cpu_setup()
        _q = _np.random.default_rng(0).uniform(-_np.pi, _np.pi, (20_000, 6))
        _t0 = time.perf_counter()
        _no.is_in_collision(_spec, _scene, _q)
        cpu_baseline["vectorised_numpy_1core_configs_per_s"] = len(_q) / (time.perf_counter() - _t0)

    import numpy as np
    import torch
    import torch.distributed as dist

    import acrobotics_b200 as ab
    from acrobotics_b200 import _native
    from acrobotics_b200.synthetic import random_configs_device, scene16

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _native.load()
    robot, scene = ab.Kuka(), scene16()
    robot_h, scene_h = robot._handle(), scene.native_handle()

    n = args.configs
    words = (n + 31) // 32
    q = random_configs_device(n, 6, seed=1000 + rank)
    # the kernel writes its mask words straight into this rank's slice of the gather buffer
    gathered = torch.zeros(world * words, dtype=torch.int32, device="cuda")
    mask = gathered[rank * words:(rank + 1) * words]
    stream = torch.cuda.current_stream().cuda_stream

    def kernel():
        _native.check(lib.acro_fk_cc_f64(robot_h, scene_h, q.data_ptr(), n, _native.Q_AOS,
                                         mask.data_ptr(), None, 0.0, stream), "acro_fk_cc_f64")

    def step(ev=None):
        if ev:
            ev[0].record()
        kernel()
        if ev:
            ev[1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _native.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        step(kev[k])
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _native.launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    kern_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / args.steps],
                           dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = world * n / (ms_per_step * 1e-3)
    w = mask[: 1 << 16].cpu().numpy().view(np.uint32)
    collision_rate = float(np.unpackbits(w.view(np.uint8)).sum()) / min(n, 32 * len(w))

    # ---- end to end: pinned host buffers through the C ABI host entry point ----
    e2e = None
    if not args.no_e2e:
        n_e2e = n
        q_host = mask_host = None
        while n_e2e >= 1 << 20:
            try:
                q_host = torch.empty((n_e2e, 6), dtype=torch.float64, pin_memory=True)
                mask_host = torch.zeros((n_e2e + 31) // 32, dtype=torch.int32, pin_memory=True)
                break
            except RuntimeError:
                n_e2e //= 2
        q_host.copy_(q[:n_e2e])
        torch.cuda.synchronize()

        def e2e_step():
            _native.check(lib.acro_fk_cc_f64_host(robot_h, scene_h, q_host.data_ptr(), n_e2e,
                                                  mask_host.data_ptr(), args.chunk),
                          "acro_fk_cc_f64_host")

        e2e_step()  # warm-up (allocates the pipeline workspace)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        same = bool((mask_host[: 1 << 12].cuda() == mask[: 1 << 12]).all().item())
        e2e = {
            "value": world * n_e2e * args.e2e_steps / dt.item(), "unit": UNIT,
            "h2d_bytes_per_step": n_e2e * 48, "d2h_bytes_per_step": ((n_e2e + 31) // 32) * 4,
            "configs_per_step_per_gpu": n_e2e, "steps": args.e2e_steps,
            "matches_device_result": same,
            "path": "acro_fk_cc_f64_host: pinned host q -> chunked H2D/K2/D2H pipeline -> host mask",
        }
        del q_host, mask_host

    if rank == 0:
        # compute roofline denominator: measured DFMA chain on this GPU
        tf = _native.C.c_double()
        pk_ms = _native.C.c_double()
        _native.check(lib.acro_measure_fma_peak(1, 4000, _native.C.byref(tf), _native.C.byref(pk_ms)))
        fp64_peak = tf.value
        _native.check(lib.acro_measure_fma_peak(0, 4000, _native.C.byref(tf), _native.C.byref(pk_ms)))
        fp32_peak = tf.value
        kms = kern_ms.item()
        achieved = W_FULL_FLOP * n / (kms * 1e-3) / 1e12
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_gbs = BYTES_PER_CONFIG * n / (kms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "configs_per_gpu_per_step": n, "scene_boxes": 16,
                "robot_boxes": 6, "self_pairs": 6, "parallelism": f"shard{world}+allgather(mask)",
                "l2": "inputs (4.8 GB/GPU) exceed the 126 MB L2; no explicit flush",
                "arithmetic": ("FP32 filter with rigorous error pads + candidate grid, exact FP64 "
                               "re-evaluation of the ambiguous configurations: verdicts "
                               "bit-identical to the all-FP64 kernel"),
                "collision_rate": round(collision_rate, 4),
            },
            "roofline": {
                # contract view: algorithmic bytes of the dominant kernel / its CUDA-event time
                "bound": "hbm", "kernel": K2_KERNEL,
                "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "traffic": K2_PROFILE.get("dram_bytes_per_config", 0) * n or None,
                "kernel_ms": kms,
                "note": ("48.125 algorithmic bytes per configuration; DRAM traffic (ncu) is 1.04x "
                         "that. The kernel is NOT HBM-bound and has no GEMM shape: it is bound by "
                         "instruction issue, see 'compute' and DESIGN.md section 4"),
                # what actually bounds the kernel
                "compute": {
                    "bound": "instruction issue (CUDA cores, FP32 filter + FP64 exact pass)",
                    "issue_slots_active_frac": (K2_PROFILE.get("issue_active_pct") or 0) / 100.0,
                    "warp_inst_per_config": K2_PROFILE.get("warp_inst_per_config"),
                    "thread_inst_per_config": K2_PROFILE.get("thread_inst_per_config"),
                    "fp32_flop_per_config": K2_PROFILE.get("fp32_flop_per_config"),
                    "fp64_flop_per_config": K2_PROFILE.get("fp64_flop_per_config"),
                    "executed_fp32_tflops": (K2_PROFILE.get("fp32_flop_per_config", 0) * n
                                             / (kms * 1e-3) / 1e12),
                    "fp32_fma_peak_measured_tflops": fp32_peak,
                    "fp64_fma_peak_measured_tflops": fp64_peak,
                    # FK + all 102 box pairs x full 15-axis SAT, no culling
                    "w_full_flop_per_config": W_FULL_FLOP,
                    "w_full_equiv_tflops": achieved,
                    "w_full_equiv_over_fp64_peak": achieved / fp64_peak,
                    "source": "ncu --set full, profiles/" + str(K2_PROFILE.get("source")),
                },
            },
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--configs", type=int, default=100_000_000, help="configurations per GPU per step")
    ap.add_argument("--chunk", type=int, default=1 << 22, help="host pipeline chunk (configs)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
