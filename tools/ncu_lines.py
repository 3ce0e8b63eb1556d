#!/usr/bin/env python
"""tools/ncu_lines.py <lib.so> <kernel-substring> <report.ncu-rep> <events> — executed warp instructions per scatter event by
SOURCE LINE, inside and outside the hot loop (nvdisasm -g line info of the library + ncu's per-instruction counts)."""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

lib, pat, rep, events = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]) / 32.0
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    text = "".join(subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout for f in glob.glob(tmp + "/*.cubin"))
on, line, sass = False, None, []
for l in text.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = pat in l
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        sass.append((m.group(2).strip(), line))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
ix = {h: i for i, h in enumerate(rows[1])}
ex = [float(r[ix["Instructions Executed"]]) for r in rows[2:]]
assert len(ex) == len(sass), (len(ex), len(sass), "the report was not made with this library")
mx = max(ex)
hot = [i for i, e in enumerate(ex) if e > 0.95 * mx]
lo, hi = hot[0], hot[-1]
inside = sum(ex[lo:hi + 1])
per_line = collections.Counter()
for i, (op, ln) in enumerate(sass):
    if not lo <= i <= hi:
        per_line[ln] += ex[i]
print(f"{sum(ex) / events:.2f} warp instructions per event: {inside / events:.2f} in the hot loop (SASS {lo}..{hi}), {sum(per_line.values()) / events:.2f} outside it:\n")
cache = {}
for (f, n), v in per_line.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 30):
    if f not in cache:
        try:
            cache[f] = open(f).read().splitlines()
        except OSError:
            cache[f] = []
    src = cache[f][n - 1].strip()[:100] if n <= len(cache[f]) else ""
    print(f"{v / events:6.3f}  {os.path.basename(f)}:{n:<4d} {src}")
