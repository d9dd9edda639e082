#!/usr/bin/env python
"""Condense `ncu --set full` reports into the per-kernel table committed under profiles/.

    python tools/profile_report.py <out.md> <rep> [<rep> ...]

For every profiled launch: duration, DRAM bytes read + written (ncu) against the ALGORITHMIC
bytes of the launch (units x bytes per unit; the unit counts are those of
tools/profile_streaming.py), achieved GB/s as algorithmic bytes / duration against the measured
copy peak of MEASURED_PEAKS.json, and the counters that say what else limits the kernel
(FP64 pipe, issue slots, occupancy, registers, top stall reasons).  Launch durations under ncu
are serialised and cold-cache: they agree with CUDA-event timings to ~5 %, the event timings
(bench.py, tools/kernel_bench.py) are the ones quoted as results.  Every launch of the capture
is listed (tools/profile_streaming.py runs each kernel twice: warm-up, then the measured one)."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 4_000_000  # tools/profile_streaming.py default
# kernel-name regex -> (label, list of (units, algorithmic bytes per unit) in launch order)
ALGO = [
    (r"fk_kernel<double, 6, 0>", "K1 fk f64", [(N, 176), (N, 176), (1_000_000, 176), (1_000_000, 176)]),
    (r"fk_kernel<float, 6, 0>", "K1 fk f32", [(N, 88)] * 2),
    (r"fk_kernel<double, 6, 1>", "K1 fk_all_links f64", [(N // 4, 816)] * 2),
    (r"fk_rpy_kernel", "K1 fk_rpy f64", [(N, 96)] * 2),
    (r"jac_kernel<6, 0>", "K3 jacobian_position", [(N, 192)] * 2),
    (r"jac_kernel<6, 1>", "K4 jacobian_rpy", [(N, 336)] * 2),
    (r"ik_kuka_kernel<0>", "K5 ik_kuka", [(N // 4, 513)] * 4 + [(2000 * 400, 513)] * 4),
    (r"path_cc_filter_kernel", "K6 path sweep (filter), segments", [(N // 8, 96.125)] * 2),
    (r"path_cc_kernel", "K6 path sweep (FP64 pass)", [(0, 0)] * 2),
    (r"fk_cc_filter_kernel", "K2 filter (over IK solutions)", [(2000 * 400 * 8, 48.125)] * 8),
    (r"fk_cc_fixup_kernel", "K2 FP64 fix-up", [(0, 0)] * 8),
    (r"csr_", "planner CSR compaction", [(0, 0)] * 16),
    (r"free_mask", "planner free mask", [(2000 * 400, 3)] * 8),
    (r"dp_persistent_kernel", "shortest path (persistent)", [(0, 0)] * 4),
    (r"dp_layer", "shortest path (per layer)", [(0, 0)] * 64),
    (r"dp_backtrack", "shortest path back-track", [(0, 0)] * 4),
]
UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def val(d, units, key):
    try:
        return float(d[key].replace(",", "")) * UNIT.get(units.get(key, ""), 1.0)
    except (KeyError, ValueError):
        return float("nan")


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except OSError:
        peak = 6650.0
    lines = ["| kernel | units | time (us) | algorithmic MB | DRAM MB (ncu) | traffic / algorithmic | "
             "GB/s (algorithmic) | frac of measured HBM | FP64 pipe % | issue % | regs | warps active % | top stalls |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    seen = {}
    for rep in reps:
        recs, units = load(rep)
        for d in recs:
            name = d.get("Kernel Name", "?")
            entry = next((e for e in ALGO if re.search(e[0], name)), None)
            label = entry[1] if entry else name[:40]
            k = seen.get(label, 0)
            seen[label] = k + 1
            n_units, bpu = entry[2][min(k, len(entry[2]) - 1)] if entry else (0, 0)
            t = val(d, units, "gpu__time_duration.sum")
            dram = val(d, units, "dram__bytes_read.sum") + val(d, units, "dram__bytes_write.sum")
            algo = n_units * bpu
            stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):
                      float(d[h]) for h in d if "issue_stalled" in h and h.endswith("per_issue_active.ratio")}
            top = ", ".join(f"{a} {b:.1f}" for a, b in sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
                            if a != "selected")
            gbs = algo / t / 1e9 if algo else float("nan")
            lines.append(
                f"| {label} | {n_units or '-'} | {t * 1e6:.1f} | {algo / 1e6:.1f} | {dram / 1e6:.1f} | "
                f"{dram / algo if algo else float('nan'):.2f} | {gbs:.0f} | {gbs / peak:.2f} | "
                f"{val(d, units, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.0f} | "
                f"{val(d, units, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
                f"{val(d, units, 'launch__registers_per_thread'):.0f} | "
                f"{val(d, units, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {top} |")
    with open(out_path, "w") as f:
        f.write(f"<!-- generated by tools/profile_report.py from {', '.join(os.path.basename(r) for r in reps)}; "
                f"HBM peak {peak} GB/s (MEASURED_PEAKS.json) -->\n")
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
