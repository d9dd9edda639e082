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


def sass_tally():
    """Executed instructions per opcode from the report's SASS source page: (warp-level counts,
    predicated-on thread-level counts).  Packed FP32 (FFMA2 / FMUL2 / FADD2) is counted here -
    the op_ffma counters above are not relied on for it."""
    import collections

    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    warp, thr = collections.Counter(), collections.Counter()
    hdr2, use, taken = None, False, False
    for row in csv.reader(txt.splitlines()):
        if not row:
            continue
        if row[0] == "Kernel Name":
            # the page may list a launch more than once: tally the first matching block only
            use = bool(pat.search(row[1])) and "fixup" not in row[1] and not taken
            taken = taken or use
            hdr2 = None
        elif row[0] == "Address":
            hdr2 = row
        elif use and hdr2 and row[0].startswith("0x"):
            src = re.sub(r"^@!?U?P\w+\s+", "", row[hdr2.index("Source")].strip())
            op = src.split()[0].split(".")[0]
            warp[op] += int(row[hdr2.index("Instructions Executed")])
            thr[op] += int(row[hdr2.index("Predicated-On Thread Instructions Executed")])
    return warp, thr


warp_ops, thr_ops = sass_tally()
fp32_flop = (2 * thr_ops["FFMA"] + thr_ops["FMUL"] + thr_ops["FADD"] + 4 * thr_ops["FFMA2"]
             + 2 * thr_ops["FMUL2"] + 2 * thr_ops["FADD2"]) / n
fp64_flop = (2 * thr_ops["DFMA"] + thr_ops["DMUL"] + thr_ops["DADD"]) / n


def source_hash():
    import hashlib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for rel in ["acrobotics_b200/csrc/fk_cc_filt.cu", "acrobotics_b200/csrc/acro_filter.cuh",
                "acrobotics_b200/csrc/acro_device.cuh", "acrobotics_b200/csrc/acro_collide.cuh",
                "acrobotics_b200/csrc/acro_math.cuh"]:
        with open(os.path.join(root, rel), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()
prof = {
    "kernel": d["Kernel Name"].split("(")[0].replace("void ", ""),
    "configs_per_launch": n,
    "gpu_time_ms": f("gpu__time_duration.sum") * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(
        rows[1][hdr.index("gpu__time_duration.sum")], 1.0),
    "thread_inst_per_config": f("sass__thread_inst_executed_true_per_opcode") / n
    if "sass__thread_inst_executed_true_per_opcode" in d else None,
    "warp_inst_per_config": f("smsp__inst_executed.sum") / n,
    "fp64_flop_per_config": fp64_flop,
    "fp32_flop_per_config": fp32_flop,
    "flop_source": "SASS page: FFMA x2, FMUL, FADD, FFMA2 x4, FMUL2 x2, FADD2 x2 (predicated-on threads)",
    "ops_per_config_ncu_counters": ops,
    "warp_inst_per_config_by_opcode": {k: v / n for k, v in warp_ops.most_common(40)},
    "arithmetic_share_of_warp_inst": sum(warp_ops[k] for k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2",
                                                                "FADD2", "DFMA", "DMUL", "DADD", "FMNMX",
                                                                "FMNMX3")) / max(1, sum(warp_ops.values())),
    "source_sha256": source_hash(),
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
