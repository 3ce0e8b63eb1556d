"""The drop-in boundary on a machine WITHOUT a GPU: the C-ABI library loads, exports every
symbol include/*.h declares, validates arguments like the header says, and refuses to compute
(no CPU fallback).  No compute call is made here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "tiny_mc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(tmc):
    lib = tmc.load()
    names = declared_symbols()
    assert len(names) >= 15 and set(names) == set(tmc.capi.EXPORTS)
    for name in names:
        assert getattr(lib, name) is not None
    dyn = subprocess.run(["nm", "-D", "--defined-only", str(tmc.lib_path())], capture_output=True, text=True, check=True).stdout
    for name in names:
        assert re.search(rf"\bT {name}\b", dyn), name
    assert lib.tmc_abi_version() == 1
    assert b"sm_100a" in lib.tmc_version()


def test_library_is_sm100a_only_and_has_no_cpu_path(tmc):
    out = subprocess.run(["cuobjdump", "-lelf", str(tmc.lib_path())], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout and "sm_90" not in out.stdout and "sm_80" not in out.stdout
    # the product never links or loads the oracle
    ldd = subprocess.run(["ldd", str(tmc.lib_path())], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libnccl" not in ldd     # NCCL is dlopen()ed only for multi-GPU init
    src = "".join(p.read_text() for p in (ROOT / "tiny_mc_b200").rglob("*") if p.suffix in {".cu", ".cuh", ".c", ".h", ".py"})
    assert "oracle" not in src.replace("oracle/README", "")


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU ") > 0
    except FileNotFoundError:
        return False


def test_fails_loudly_without_a_gpu(tmc):
    if _has_gpu():
        pytest.skip("a GPU is present; covered by the gpu tests")
    with pytest.raises(tmc.TinyMcError) as e:
        tmc.init(1)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)
    heat = np.zeros(101, np.float32)
    with pytest.raises(tmc.TinyMcError) as e:
        tmc.photons("default", 1, 0, 10, heat, heat.copy())
    assert e.value.code == 1
    assert not heat.any()


def test_argument_validation(tmc):
    lib = tmc.load()
    s = tmc.Scales()
    bad = [tmc.Params(0, 2.0, 20.0, 50.0), tmc.Params(101, 0.0, 20.0, 50.0), tmc.Params(101, 2.0, -1.0, 50.0),
           tmc.Params(101, 2.0, 20.0, 0.0)]
    bad.append(tmc.Params(101, 1e-7, 100.0, 50.0))       # absorbed fraction 1e-9 per event: > 2^22 events per generation
    for p in bad:
        assert lib.tmc_fx_scales(C.byref(p), C.byref(s)) == 2
        assert lib.tmc_last_error()
    assert lib.tmc_fx_scales(None, C.byref(s)) == 2
    assert lib.tmc_set_option(b"philox_rounds", 8) == 2
    assert lib.tmc_set_option(b"no_such_option", 1) == 2
    assert lib.tmc_set_option(b"philox_rounds", 10) == 0
    for name, good, bad_value in ((b"batch_streams", 1, 3), (b"batch_capacity", 64, 5000), (b"tally_layout", 1, 7)):
        assert lib.tmc_set_option(name, bad_value) == 2 and lib.tmc_set_option(name, good) == 0 and lib.tmc_set_option(name, 0) == 0
    assert lib.tmc_last_run_info(None) == 2
    # the batched and the device-checking entry points validate before they touch a device
    p_ok = tmc.capi.make_params("default")
    h = np.zeros(101, np.uint64)
    assert lib.tmc_photons_fx_batches(C.byref(p_ok), 1, 0, 10, 1, None, h.ctypes.data) in (1, 2)     # not initialised / NULL
    assert lib.tmc_device_tallies_check(None, 0, None, None) == 2


@pytest.mark.parametrize("name", ["default", "highalbedo", "finegrid"])
def test_fixed_point_plan_matches_oracle_restatement(tmc, orc, name):
    """Product (tmc_fx_scales) and oracle (orc_fx_plan) derive the same scales independently."""
    a, b = tmc.fx_scales(name), orc.fx_plan(name)
    assert (a.heat_shift, a.heat2_rshift, a.absorb_q32, a.roulette_thr) == (b.heat_shift, b.heat2_rshift, b.absorb_q32, b.roulette_thr)
    cfg = tmc.CONFIGS[name]
    albedo = np.float32(cfg["mu_s"]) / (np.float32(cfg["mu_s"]) + np.float32(cfg["mu_a"]))   # reference photon.c:8
    assert abs(a.absorb_q32 / 2.0**32 - (1.0 - float(albedo))) < 2.0**-32
    dep_max = (1 << a.heat_shift) * a.absorb_q32 >> 32
    assert 2**16 <= dep_max < 2**18
    assert abs(a.roulette_thr / 2.0**a.heat_shift - 0.001) < 1e-6                            # reference photon.c:45


def test_fx_accumulate_adds_like_photon_does(tmc):
    """photon() only ever ADDS into the caller's arrays (reference photon.c:30-31)."""
    sc = tmc.fx_scales("default")
    heat_fx = np.arange(101, dtype=np.uint64) << np.uint64(sc.heat_shift - 3)
    heat2_fx = np.arange(101, dtype=np.uint64) * np.uint64(3)
    heats = np.full(101, 1.5, np.float32)
    heats2 = np.full(101, 0.25, np.float32)
    tmc.fx_accumulate("default", heat_fx, heat2_fx, heats, heats2)
    assert np.allclose(heats, 1.5 + np.arange(101) / 8.0)
    assert np.allclose(heats2, 0.25 + 3 * np.arange(101) * 2.0 ** (sc.heat2_rshift - 2 * sc.heat_shift), rtol=1e-6)


@pytest.mark.parametrize("name", ["default", "highalbedo", "finegrid"])
def test_weight_schedule_matches_oracle_and_the_reference_walk(tmc, orc, name):
    """The deterministic weight schedule the kernel is built on: product and
    oracle derive it independently, and it is what the reference's float arithmetic does
    (reference photon.c:32,45-48: w *= albedo until w < 0.001, then x10 on survival)."""
    mine = tmc.generation_plan(name, 8)
    theirs = orc.generation_plan(name, 8)
    for a, b in zip(mine, theirs):
        assert np.array_equal(a, b)
    first, n, w = mine
    assert first[0] == 1 and np.array_equal(first[1:], first[:-1] + n[:-1])
    cfg = tmc.CONFIGS[name]
    albedo = np.float32(cfg["mu_s"]) / (np.float32(cfg["mu_s"]) + np.float32(cfg["mu_a"]))
    wf, k = np.float32(1.0), 0                      # generation 0 in the reference's own float arithmetic
    while True:
        wf = np.float32(wf * albedo)
        k += 1
        if wf < np.float32(0.001):
            break
    assert n[0] == k                                # 73 (default, fine grid), 6912 (high albedo): SURVEY §4
    assert abs(float(w[1]) / float(w[0]) - 10.0 * float(wf)) < 2e-5
