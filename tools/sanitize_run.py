#!/usr/bin/env python
"""tools/sanitize_run.py — a small walk of every configuration, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiny_mc_b200 as tmc  # noqa: E402

tmc.init(1)
for name, n in (("default", 20000), ("highalbedo", 200), ("finegrid", 20000)):
    h, h2 = tmc.photons_fx(name, 7, 123, n)
    info = tmc.last_run_info()
    print(name, n, info.events, int(h.sum()), flush=True)
tmc.finalize()
