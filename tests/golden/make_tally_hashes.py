#!/usr/bin/env python
"""tests/golden/make_tally_hashes.py — run on a B200: the FNV-1a-64 hashes of the kernel's exact tally words
(heat_fx | heat2_fx, u64) for fixed photon ranges of the three configurations -> tests/golden/tally_hashes.json.

The tallies are integers and a photon's trajectory depends on (seed, photon index) only, so these words are a
pure function of the kernel's arithmetic: any change of the stream, the tables, the rounding of one FFMA or the
shell index of one event changes them.  Regenerate ONLY together with a deliberate change of the stream or the
arithmetic (and say so in DESIGN.md); an optimisation that keeps the arithmetic must reproduce them
(tests/test_gpu_parity.py::test_tallies_are_the_committed_words)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import tiny_mc_b200 as tmc   # noqa: E402
from stats import fnv64      # noqa: E402

CASES = {"default": (0x5EED, 0, 1 << 26), "highalbedo": (0x5EED, 0, 1 << 18), "finegrid": (0x5EED, 0, 1 << 24),
         "default_far_range": (24301, (1 << 40) + 12345, 1 << 22)}

tmc.init(1)
out = {}
for name, (seed, first, n) in CASES.items():
    h, h2 = tmc.photons_fx(name.split("_")[0], seed, first, n)
    out[name] = {"seed": seed, "first": first, "photons": n, "fnv1a64": fnv64(np.concatenate([h, h2])), "events": int(tmc.last_run_info().events)}
tmc.finalize()
(Path(__file__).resolve().parent / "tally_hashes.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))
