#!/usr/bin/env python
"""Per-source-line and per-opcode tallies of executed instructions for one kernel.

    python tools/sass_by_line.py <ncu source-page csv> <object file> <mangled-name substring>
                                 [--top 40] [--configs N]

Joins `ncu -i X.ncu-rep --page source --csv` (SASS view: one row per instruction with
"Instructions Executed" / "Thread Instructions Executed") with the line table of the same
kernel in the object file (`nvdisasm -g`), by instruction offset.  The object must be the
build the report was captured from (checked: opcodes must agree).
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, name_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout.splitlines()
    rows, inside, cur = [], False, None
    for ln in txt:
        if ln.startswith("\t.section\t.text."):
            inside = name_sub in ln
            cur = None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return rows


def opcode(s):
    s = re.sub(r"^@!?U?P\d+\s+", "", s.strip())
    return s.split()[0].split(".")[0] if s else "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("obj")
    ap.add_argument("name")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--configs", type=float, default=0.0, help="units per launch (per-unit columns)")
    ap.add_argument("--kernel-index", type=int, default=0, help="which launch of the csv")
    args = ap.parse_args()

    table = line_table(args.obj, args.name)
    launches, cur = [], None
    with open(args.csv, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "rows": [], "hdr": None}
                launches.append(cur)
            elif row[0] == "Address" and cur is not None:
                cur["hdr"] = row
            elif cur is not None and cur["hdr"] is not None and row[0].startswith("0x"):
                cur["rows"].append(row)
    sel = [l for l in launches if args.name in l["name"] or True]
    L = sel[args.kernel_index]
    hdr = L["hdr"]
    iS, iI, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index(
        "Thread Instructions Executed")
    if len(L["rows"]) != len(table):
        print(f"warning: {len(L['rows'])} profiled instructions vs {len(table)} in the object",
              file=sys.stderr)
    by_line = collections.Counter()
    by_line_t = collections.Counter()
    by_op = collections.Counter()
    by_file = collections.Counter()
    mism = 0
    tot = tot_t = 0
    for k, row in enumerate(L["rows"]):
        inst, thr = int(row[iI]), int(row[iT])
        op = opcode(row[iS])
        loc = None
        if k < len(table):
            if opcode(table[k][1]) != op:
                mism += 1
            loc = table[k][2]
        by_line[loc] += inst
        by_line_t[loc] += thr
        by_op[op] += inst
        by_file[loc[0] if loc else None] += inst
        tot += inst
        tot_t += thr
    if mism:
        print(f"warning: {mism} opcode mismatches - object is not the profiled build", file=sys.stderr)
    per = (lambda v: f"{v / args.configs:8.2f}") if args.configs else (lambda v: "")
    print(f"kernel: {L['name'][:100]}")
    print(f"warp instructions {tot}  thread instructions {tot_t}  lanes/inst {tot_t / max(tot, 1):.2f}")
    print("\n-- opcode mix (warp instructions)")
    for op, v in by_op.most_common(30):
        print(f"{op:12s} {100.0 * v / tot:6.2f} % {per(v)}")
    print("\n-- by source line")
    for loc, v in by_line.most_common(args.top):
        lanes = by_line_t[loc] / max(v, 1)
        print(f"{str(loc):40s} {100.0 * v / tot:6.2f} % lanes {lanes:5.1f} {per(v)}")
    print("\n-- by file")
    for fn, v in by_file.most_common():
        print(f"{str(fn):30s} {100.0 * v / tot:6.2f} %")


if __name__ == "__main__":
    main()
