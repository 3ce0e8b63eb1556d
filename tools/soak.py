#!/usr/bin/env python
"""tools/soak.py — very large runs through the C ABI (launch splitting at 2^30 / 2^32 boundaries, u64 tallies):
size-independent properties only."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tiny_mc_b200 as tmc  # noqa: E402

tmc.init(1)
out = []
for name, first, n in (("default", (1 << 33) - 12345, 1 << 36), ("highalbedo", 7, 1 << 31), ("finegrid", 1 << 40, 1 << 34)):
    t0 = time.perf_counter()
    hfx, h2fx = tmc.photons_fx(name, 20261017, first, n)
    dt = time.perf_counter() - t0
    info = tmc.last_run_info()
    heat, heat2 = tmc.capi.fx_to_float64(name, hfx, h2fx)
    cfg = tmc.CONFIGS[name]
    a = float(np.float32(cfg["mu_s"]) / (np.float32(cfg["mu_s"]) + np.float32(cfg["mu_a"])))
    out.append(dict(config=name, first=first, photons=n, seconds=dt, photons_per_s=n / dt, launches=info.gpu_launches, retries=info.retries,
                    absorbed_per_photon_minus_1=float(heat.sum() / n - 1.0), heat2_ratio_to_closed_form=float(heat2.sum() / n / ((1 - a) / (1 + a))),
                    events_per_photon=info.events / n, counted_photons=int(info.photons), extra=float(heat[-1] / n)))
    print(json.dumps(out[-1]), flush=True)
tmc.finalize()
