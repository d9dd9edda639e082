#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per block of SASS instructions the
stall samples, executed warp instructions and average active threads."""
import csv
import sys

path, block = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
# the CSV holds one segment per profiled launch: "Kernel Name" row, header row, instructions
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo_seg = starts[which]
hi_seg = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[lo_seg][1][:100])
hdr = rows[lo_seg + 1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[lo_seg + 2:hi_seg]
tot_s = sum(float(r[col["# Samples"]] or 0) for r in data)
tot_i = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e}")
for lo in range(0, len(data), block):
    blk = data[lo:lo + block]
    s = sum(float(r[col["# Samples"]] or 0) for r in blk)
    i = sum(float(r[col["Instructions Executed"]] or 0) for r in blk)
    t = sum(float(r[col["Thread Instructions Executed"]] or 0) for r in blk)
    lsb = sum(float(r[col["stall_long_sb"]] or 0) for r in blk)
    ops = {}
    for r in blk:
        op = r[col["Source"]].split()[0] if r[col["Source"]] else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = ",".join(f"{k}{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:5])
    print(f"{lo:5d} samples {100*s/tot_s:5.1f}%  inst {100*i/tot_i:5.1f}%  thr/inst {t/max(i,1):5.1f}  long_sb {100*lsb/max(s,1):4.0f}%  {top}")
