#!/usr/bin/env python
"""tools/quick_bench.py — kernel-time photons/s for the named configs (library CUDA events), one line each."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiny_mc_b200 as tmc  # noqa: E402

tmc.init(1)
plans = [("default", 1 << 26, 10, 0, 0), ("default", 1 << 26, 7, 0, 0), ("default", 1 << 26, 10, 256, 2), ("default", 1 << 26, 10, 256, 3),
         ("default", 1 << 26, 10, 256, 4), ("default", 1 << 26, 10, 512, 1), ("default", 1 << 26, 10, 512, 2), ("default", 1 << 26, 10, 1024, 1),
         ("default", 1 << 26, 10, 128, 4), ("highalbedo", 1 << 20, 10, 0, 0), ("highalbedo", 1 << 20, 10, 1024, 1), ("highalbedo", 1 << 20, 10, 256, 2),
         ("highalbedo", 1 << 20, 10, 256, 3), ("highalbedo", 1 << 20, 10, 128, 4), ("finegrid", 1 << 26, 10, 0, 0), ("finegrid", 1 << 26, 10, 512, 1)]
if len(sys.argv) > 1:
    plans = [p for p in plans if p[0] in sys.argv[1:]]
for name, n, rounds, block, per_sm in plans:
    tmc.set_option("philox_rounds", rounds)
    tmc.set_option("block_threads", block)
    tmc.set_option("blocks_per_sm", per_sm)
    tmc.photons_fx(name, 1, 0, n >> 3)
    best = None
    for rep in range(3):
        tmc.photons_fx(name, 1, rep * n, n)
        info = tmc.last_run_info().as_dict()
        if best is None or info["kernel_ms"] < best["kernel_ms"]:
            best = info
    print(json.dumps(dict(config=name, rounds=rounds, photons_per_s=n / best["kernel_ms"] * 1e3, events_per_s=best["events"] / best["kernel_ms"] * 1e3,
                          kernel_ms=best["kernel_ms"], block=best["threads_per_block"], grid=best["blocks_per_gpu"], flush=best["flush_iters"])), flush=True)
tmc.finalize()
