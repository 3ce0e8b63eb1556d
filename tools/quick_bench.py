#!/usr/bin/env python
"""tools/quick_bench.py — kernel-time photons/s (library CUDA events), one JSON line per plan.

    python tools/quick_bench.py [config:block:per_sm[:rounds[:log2n]] ...]     (TMC_LIB selects a variant library)
"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiny_mc_b200 as tmc  # noqa: E402

DEFAULT_N = {"default": 26, "highalbedo": 20, "finegrid": 26}
plans = sys.argv[1:] or ["default:0:0", "default:512:1", "default:768:1", "default:1024:1", "default:256:2", "default:256:3",
                         "highalbedo:0:0", "highalbedo:768:1", "highalbedo:1024:1", "finegrid:0:0", "finegrid:512:1", "finegrid:768:1"]
tmc.init(1)
for plan in plans:
    f = plan.split(":")
    name, block, per_sm = f[0], int(f[1]), int(f[2])
    rounds = int(f[3]) if len(f) > 3 else 10
    n = 1 << (int(f[4]) if len(f) > 4 else DEFAULT_N[name])
    try:
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
        print(json.dumps(dict(lib=os.path.basename(os.environ.get("TMC_LIB", "")), config=name, rounds=rounds,
                              photons_per_s=n / best["kernel_ms"] * 1e3, events_per_s=best["events"] / best["kernel_ms"] * 1e3,
                              kernel_ms=best["kernel_ms"], block=best["threads_per_block"], grid=best["blocks_per_gpu"],
                              flush=best["flush_iters"], smem=best["smem_bytes"])), flush=True)
    except tmc.TinyMcError as e:
        print(json.dumps(dict(plan=plan, error=str(e))), flush=True)
tmc.finalize()
