#!/usr/bin/env python
"""tools/ncu_stalls.py — per-opcode share of executed instructions and of warp-stall samples, with the
stall reasons, from the source page of an .ncu-rep (needs -lineinfo and --import-source on).

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep > profiles/xyz_stalls.md
"""
import collections
import csv
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
print(f"# Warp-stall samples by opcode — `{rows[0][1]}`\n")
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


reasons = ["stall_wait", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_math", "stall_mio", "stall_short_sb",
           "stall_long_sb", "stall_branch_resolving", "stall_no_inst", "stall_barrier"]
tot = sum(f(r, "# Samples") for r in data)
allr = collections.Counter()
agg, ex, by = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for r in data:
    src = r[ix["Source"]].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0] if src else "?"
    base = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op if op.startswith("MUFU") else op.split(".")[0]
    agg[base] += f(r, "# Samples")
    ex[base] += f(r, "Instructions Executed")
    for k in reasons:
        by[base][k] += f(r, k)
        allr[k] += f(r, k)
print("all samples: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f} %" for k, v in allr.most_common()) + "\n")
totex = sum(ex.values())
print("| opcode | executed % | samples % | " + " | ".join(k[6:] for k in reasons[:7]) + " |")
print("|---|---:|---:|" + "---:|" * 7)
for op, v in agg.most_common(18):
    n = max(v, 1.0)
    print(f"| {op} | {100 * ex[op] / totex:.1f} | {100 * v / tot:.1f} | " + " | ".join(f"{100 * by[op][k] / n:.0f}" for k in reasons[:7]) + " |")
print("\n(reason columns: share of that opcode's samples; `selected` = the warp was issuing)")
