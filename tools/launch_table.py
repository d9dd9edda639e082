#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
markdown table (per kernel: launches, average ms, share).

    python tools/launch_table.py launches.csv "command that was profiled" [plain ms/step] > profiles/x.md
"""
import collections
import csv
import sys

path, cmd = sys.argv[1], sys.argv[2]
plain_ms = float(sys.argv[3]) if len(sys.argv) > 3 else None
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
i_grid, i_blk = hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[i_name]
    if "distribution" in name:
        short = "at::native::uniform_ (torch RNG: synthetic input generation)"
    elif "at::" in name:
        short = "at::native::" + name.split("<")[0].split("::")[-1] + " (torch: input generation / bookkeeping)"
    else:
        short = name.split("(")[0].replace("void ", "")
    agg.setdefault((short, r[i_grid], r[i_blk]), []).append(float(r[i_val]) / 1e6)
tot = sum(sum(v) for v in agg.values())
print(f"# ncu launch list\n\nCommand: `{cmd}`  \nCollected with `ncu --metrics gpu__time_duration.sum "
      "--clock-control none --csv` on one B200.  Per-launch times are cold-cache and serialised: "
      "compare shares, not absolutes.\n")
print("| kernel | grid | block | launches | avg ms | share of all launches |\n|---|---|---|---|---|---|")
for (n, g, b), v in agg.items():
    print(f"| `{n}` | {g} | {b} | {len(v)} | {sum(v)/len(v):.4f} | {100*sum(v)/tot:.2f}% |")
ours = {k: v for k, v in agg.items() if k[0].startswith("acro::")}
print("\nKernels of libacro_b200.so in this run:")
for (n, g, b), v in ours.items():
    print(f"* `{n}`: {len(v)} launches, {sum(v)/len(v):.4f} ms each under ncu")
if plain_ms is not None:
    print(f"\nPlain run (no profiler), CUDA events: {plain_ms:.3f} ms per step.")
