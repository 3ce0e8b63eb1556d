#!/usr/bin/env python
"""tools/sass_sched.py <lib.so> <kernel-substring> [--dump] — static schedule of a kernel's hottest loop:
instructions, sum of the control-code stall counts (the cycles ONE warp needs to issue the loop once, before any
scoreboard wait), and how far the schedule interleaves independent work (average stall per instruction; 1 = ideal)."""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
dump = "--dump" in sys.argv
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
on = False
lines = []
for l in out.splitlines():
    if "Function :" in l:
        on = pat in l
    if on:
        lines.append(l)
ins = []
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        if m2:
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
            i += 2
            continue
    i += 1
loops = []
for k, (addr, op, w1) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", op)
    if m:
        tgt = int(m.group(1), 16) // 16
        if tgt <= k:
            loops.append((tgt, k))
best = None
for s, e in loops:
    n_mufu = sum("MUFU" in ins[k][1] for k in range(s, e + 1))
    if n_mufu >= 32 and (best is None or e - s < best[1] - best[0]):
        best = (s, e, n_mufu)
s, e, n_mufu = best
tot = 0
hist = collections.defaultdict(lambda: [0, 0])
for k in range(s, e + 1):
    addr, op, w1 = ins[k]
    stall = (w1 >> 41) & 0xF
    tot += stall
    name = re.sub(r"^@!?U?P\d+\s+", "", op).split()[0]
    hist[name][0] += 1
    hist[name][1] += stall
    if dump:
        wr, rd, wait = (w1 >> 46) & 7, (w1 >> 49) & 7, (w1 >> 52) & 0x3F
        print(f"{k:5d} s{stall:2d} w{wr if wr != 7 else '-'} r{rd if rd != 7 else '-'} m{wait:02x}  {op}")
n = e - s + 1
print(f"kernel {pat}: {len(ins)} instructions; hot loop {s}..{e}: {n} instructions, {n_mufu} MUFU (= {n_mufu // 2} events), "
      f"stall-count sum {tot} cycles = {tot / n:.2f} per instruction, {tot / (n_mufu / 2):.1f} cycles per event for one warp")
for k, v in sorted(hist.items(), key=lambda x: -x[1][0]):
    print(f"  {k:18s} n={v[0]:3d} stall_sum={v[1]:4d} avg={v[1] / v[0]:.2f}")
