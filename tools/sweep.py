#!/usr/bin/env python
"""tools/sweep.py — launch-shape sweep of the walk kernel on one B200 (run under gpurun).

Prints one JSON line per (config, block_threads, blocks_per_sm, flush_iters, philox_rounds):
photons/s from the library's own CUDA-event kernel time (tmc_run_info.kernel_ms).
"""
import itertools
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiny_mc_b200 as tmc  # noqa: E402

PLANS = {
    "default": dict(n=1 << 25, blocks=[128, 256, 512], per_sm=[0, 1, 2], flush=[0, 8], rounds=[10, 7]),
    "highalbedo": dict(n=1 << 19, blocks=[256, 512], per_sm=[0, 2], flush=[0], rounds=[10]),
    "finegrid": dict(n=1 << 25, blocks=[256, 512], per_sm=[0], flush=[0, 4], rounds=[10]),
}


def main():
    names = sys.argv[1:] or list(PLANS)
    tmc.init(1)
    for name in names:
        plan = PLANS[name]
        tmc.photons_fx(name, 1, 0, 1 << 20)   # warm-up
        for block, per_sm, flush, rounds in itertools.product(plan["blocks"], plan["per_sm"], plan["flush"], plan["rounds"]):
            if per_sm and block * per_sm > 2048:
                continue
            tmc.set_option("block_threads", block)
            tmc.set_option("blocks_per_sm", per_sm)
            tmc.set_option("flush_iters", flush)
            tmc.set_option("philox_rounds", rounds)
            try:
                best = None
                for rep in range(2):
                    tmc.photons_fx(name, 1, rep * plan["n"], plan["n"])
                    info = tmc.last_run_info().as_dict()
                    if best is None or info["kernel_ms"] < best["kernel_ms"]:
                        best = info
                best.update(config=name, photons_per_s=plan["n"] / (best["kernel_ms"] * 1e-3),
                            events_per_s=best["events"] / (best["kernel_ms"] * 1e-3), blocks_per_sm_opt=per_sm)
                print(json.dumps(best), flush=True)
            except tmc.TinyMcError as e:
                print(json.dumps(dict(config=name, block=block, per_sm=per_sm, flush=flush, rounds=rounds, error=str(e))), flush=True)
    tmc.finalize()


if __name__ == "__main__":
    main()
