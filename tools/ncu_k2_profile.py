#!/usr/bin/env python
"""Extract the per-configuration counters of K2 from an `ncu --set full` report
and write profiles/k2_profile.json (read by bench.py for roofline.executed / traffic).

    python tools/ncu_k2_profile.py gpurun_out/prof.ncu-rep <configs per launch> [kernel regex]
"""
import csv
import json
import os
import re
import subprocess
import sys

rep, n = sys.argv[1], int(sys.argv[2])
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else re.compile("fk_cc")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
recs = [dict(zip(hdr, r)) for r in rows[2:] if pat.search(dict(zip(hdr, r)).get("Kernel Name", ""))]
d = recs[-1]


def f(key, default=0.0):
    try:
        return float(d[key].replace(",", ""))
    except (KeyError, ValueError):
        return default


cycles = f("sm__cycles_elapsed.avg")
n_smsp = f("smsp__cycles_elapsed.sum") / max(cycles, 1) if "smsp__cycles_elapsed.sum" in d else 592


def per_config(op):
    """thread-level executed ops of one SASS class per configuration"""
    k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
    if k in d:
        return f(k) / n
    k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
    return f(k) * cycles / n


ops = {op: per_config(op) for op in ("dfma", "dmul", "dadd", "ffma", "fmul", "fadd")}
prof = {
    "kernel": d["Kernel Name"].split("(")[0].replace("void ", ""),
    "configs_per_launch": n,
    "gpu_time_ms": f("gpu__time_duration.sum") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(
        rows[1][hdr.index("gpu__time_duration.sum")], 1.0),
    "thread_inst_per_config": f("sass__thread_inst_executed_true_per_opcode") / n
    if "sass__thread_inst_executed_true_per_opcode" in d else None,
    "warp_inst_per_config": f("smsp__inst_executed.sum") / n,
    "fp64_flop_per_config": 2 * ops["dfma"] + ops["dmul"] + ops["dadd"],
    "fp32_flop_per_config": 2 * ops["ffma"] + ops["fmul"] + ops["fadd"],
    "ops_per_config": ops,
    "dram_bytes_per_config": (f("dram__bytes_read.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(rows[1][hdr.index("dram__bytes_read.sum")], 1)
                              + f("dram__bytes_write.sum") * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(rows[1][hdr.index("dram__bytes_write.sum")], 1)) / n,
    "registers_per_thread": f("launch__registers_per_thread"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "source": os.path.basename(rep),
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k2_profile.json")
json.dump(prof, open(path, "w"), indent=1)
print(json.dumps(prof, indent=1))
