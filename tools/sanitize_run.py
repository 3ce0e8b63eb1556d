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
bh, bh2 = tmc.photons_fx_batches("default", 7, 123, 20000, 5)       # slot-major launches on two streams
h, h2 = tmc.photons_fx("default", 7, 123, 20000)
assert (bh.sum(axis=0) == h).all() and (bh2.sum(axis=0) == h2).all()
print("batches", int(bh.sum()), flush=True)
tmc.finalize()
