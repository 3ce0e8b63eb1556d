"""bench.py's reference arm runs without a GPU: its stdout must be exactly one JSON line with the
keys the driver reads (the GPU arm prints the same keys plus roofline; checked on the GPU box)."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-photons-per-core", "512"], capture_output=True, text=True, check=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "photons/s" and d["unit"] == "photons/s" and d["higher_is_better"] is True
    assert d["value"] > 1e4 and d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


import pytest


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_keys():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, check=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["metric"] == "photons/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["gpu_launches"] == 2 and d["vs_baseline"] is None
    assert d["value"] > 1e9 and 0.5 < d["e2e"]["value"] / d["value"] < 1.1 and d["e2e"]["d2h_bytes_per_step"] == 8 * (2 * 101 + 4)
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac_walk_only_35_slots"] < 1.0
    # frac is the hardware issue-slot utilisation (live events/s x executed instructions per event of the committed
    # ncu capture), not the canonical-budget figure: below 1 by construction, and below the dispatch-slot figure
    assert 0.3 < r["frac"] < 1.0 and r["frac"] < r["frac_dispatch"] < 1.0 and r["frac_canonical_88"] > r["frac"]
    assert r["traffic"] is not None and r["ncu_capture"]["matches_this_run"]
    assert d["scaling"] == "weak" and d["also"]["photons_per_step"] == 1 << 29 and d["also"]["value"] > 1e9
    # the fixed-range tally hash: the torch-stream path and the library's own host path give the same words
    c = d["checks"]
    assert len(c["tally_hash"]) == 16 and c["tally_hash"] == c["tally_hash_library_path"]
    assert "workload" in d["config"] and "BASELINE configs[1]" in d["config"]["workload"]
    assert d["checks"]["tally_range_flag"] == 0 and abs(d["checks"]["absorbed_weight_per_photon"] - 1.0) < 1e-4
