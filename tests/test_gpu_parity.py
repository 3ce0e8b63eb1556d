"""Parity of the CUDA walk (through the C ABI) with the oracle — needs a B200 (`-m gpu`).

Three layers, strongest first:
 1. replay: the oracle's CPU replay of the product's own Philox stream (oracle/stream_replay.c)
    must give the SAME integers — scatter events, total fixed-point weight, total squared
    deposits — and the same per-shell tallies up to shell flips where MUFU and libm round a
    radius to opposite sides of a shell boundary;
 2. reproducibility: the tally words do not depend on the split of the photon range, the block
    shape, the drain interval or the number of GPUs (bit-identical);
 3. statistics vs the reference walk (reference photon.c:6-51): per-shell means within 4 sigma
    (batch-means sigma), total absorbed weight to 1e-4 relative, against the
    committed fixtures made from the UNMODIFIED reference object code and from its bit-exact
    port on a sound generator (tests/golden/, oracle/make_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from conftest import GOLDEN, ROOT
from stats import batch_means_z, fnv64, literal_sigma_z

pytestmark = pytest.mark.gpu

SEED = 0x5EED


def albedo(cfg):
    return float(np.float32(cfg["mu_s"]) / (np.float32(cfg["mu_s"]) + np.float32(cfg["mu_a"])))   # reference photon.c:8


# ------------------------------------------------------------------ 1. replay (integer-exact)
@pytest.mark.parametrize("name,n,flip_tol", [("default", 20000, 4e-4), ("highalbedo", 150, 4e-4), ("finegrid", 20000, 6e-3)])
def test_gpu_equals_cpu_replay_of_the_same_stream(gpu, orc, name, n, flip_tol):
    heat_fx, heat2_fx = gpu.photons_fx(name, SEED, 1000, n)
    info = gpu.last_run_info()
    r_heat, r_heat2, r_events = orc.replay(name, SEED, 1000, n)
    assert info.photons == n
    assert info.events == r_events                          # every photon lives exactly as long
    assert int(heat_fx.sum()) == int(r_heat.sum())           # every deposit is the same integer
    assert int(heat2_fx.sum()) == int(r_heat2.sum())
    moved = np.abs(heat_fx.astype(np.int64) - r_heat.astype(np.int64)).sum() / 2
    assert moved <= flip_tol * r_heat.sum(), f"{moved / r_heat.sum():.2e} of the weight changed shell (MUFU vs libm)"
    assert info.retries == 0


def test_tallies_are_the_committed_words(gpu):
    """The exact tally words of fixed photon ranges, hashed (tests/golden/make_tally_hashes.py): an optimisation of
    the kernel that keeps its arithmetic (scheduling, packed f32x2 forms, launch shapes, queue layout ...) must
    reproduce them bit for bit; the default-optics entry is bench.py's checks.tally_hash."""
    import json
    golden = json.loads((GOLDEN / "tally_hashes.json").read_text())
    for name, g in golden.items():
        h, h2 = gpu.photons_fx(name.split("_")[0], g["seed"], g["first"], g["photons"])
        assert gpu.last_run_info().events == g["events"], name
        assert fnv64(np.concatenate([h, h2])) == g["fnv1a64"], name


@pytest.mark.parametrize("rounds", [7, 10])
def test_philox_round_variants_match_replay(gpu, orc, rounds):
    gpu.set_option("philox_rounds", rounds)
    try:
        heat_fx, heat2_fx = gpu.photons_fx("default", 99, 0, 3000)
        ev = gpu.last_run_info().events
    finally:
        gpu.set_option("philox_rounds", 10)
    r_heat, r_heat2, r_events = orc.replay("default", 99, 0, 3000, rounds=rounds)
    assert ev == r_events and int(heat_fx.sum()) == int(r_heat.sum()) and int(heat2_fx.sum()) == int(r_heat2.sum())


def test_philox7_and_philox10_tallies_are_statistically_identical(gpu):
    """The documented statistical test behind the optional 7-round generator: at
    2^30 photons per side (per-shell precision ~3e-5) every shell of the Philox4x32-7 walk is
    within 4.5 sigma of the Philox4x32-10 walk, and so is the total absorbed weight."""
    nb, n = 32, 1 << 25
    sides = {}
    for rounds in (10, 7):
        gpu.set_option("philox_rounds", rounds)
        try:
            sides[rounds] = gpu_batches(gpu, "default", nb, n, seed=2024 + rounds)
        finally:
            gpu.set_option("philox_rounds", 10)
    z, ok = batch_means_z(sides[7][0], n, sides[10][0], n)
    assert ok.all() and np.abs(z).max() < 4.5, z
    assert abs(z.mean()) < 0.6
    tot = [sides[r][0].sum() / (nb * n) for r in (7, 10)]
    assert abs(tot[0] - tot[1]) < 6 * 0.00301 / np.sqrt(nb * n) * np.sqrt(2)


def test_single_photon_and_empty_range(gpu, orc):
    h, h2 = gpu.photons_fx("default", 5, 12345678901, 1)
    r, r2, ev = orc.replay("default", 5, 12345678901, 1)       # photon index > 2^32: high counter word
    assert gpu.last_run_info().events == ev and int(h.sum()) == int(r.sum())
    assert np.abs(h.astype(np.int64) - r.astype(np.int64)).sum() <= 2 * r.max()
    # a range that crosses a multiple of 2^32 is cut into two launches (32-bit photon offsets)
    lo = (1 << 32) - 1000
    h, h2 = gpu.photons_fx("default", 5, lo, 3000)
    r, r2, ev = orc.replay("default", 5, lo, 3000)
    assert gpu.last_run_info().events == ev and int(h.sum()) == int(r.sum()) and gpu.last_run_info().gpu_launches == 2
    h, h2 = gpu.photons_fx("default", 5, 0, 0)
    assert not h.any() and not h2.any() and gpu.last_run_info().photons == 0


@pytest.mark.parametrize("cfg,n", [
    (dict(shells=101, mu_a=5.0, mu_s=0.0, microns_per_shell=50.0), 20000),        # pure absorber: one event per generation
    (dict(shells=1, mu_a=2.0, mu_s=20.0, microns_per_shell=50.0), 5000),          # a single (overflow) bin
    (dict(shells=37, mu_a=1.0, mu_s=3.0, microns_per_shell=400.0), 20000),        # low albedo, coarse shells: generations of 5 events
    (dict(shells=512, mu_a=0.5, mu_s=60.0, microns_per_shell=20.0), 4000),        # largest lane-private grid
    (dict(shells=513, mu_a=0.5, mu_s=60.0, microns_per_shell=20.0), 4000),        # smallest shared-histogram grid
    (dict(shells=2000, mu_a=3.0, mu_s=9.0, microns_per_shell=500.0), 20000),      # coarse shells in the shared histogram: hot bins
])
def test_other_optics_equal_the_replay(gpu, orc, cfg, n):
    """Run-time parameters (reference params.h:5-23 are compile-time): whatever the optics, events
    and fixed-point totals equal the CPU replay exactly, per-shell words up to MUFU shell flips."""
    heat_fx, heat2_fx = gpu.photons_fx(cfg, 31337, 5, n)
    info = gpu.last_run_info()
    r_heat, r_heat2, r_events = orc.replay(cfg, 31337, 5, n)
    assert info.events == r_events and info.retries == 0
    assert int(heat_fx.sum()) == int(r_heat.sum()) and int(heat2_fx.sum()) == int(r_heat2.sum())
    moved = np.abs(heat_fx.astype(np.int64) - r_heat.astype(np.int64)).sum() / 2
    assert moved <= 2e-3 * r_heat.sum()
    if cfg["mu_s"] == 0.0:
        # every photon deposits its whole weight at the first collision: heat[s] / N is the exponential
        # step distribution integrated over the shell (reference photon.c:21,26)
        heat, _ = gpu.capi.fx_to_float64(cfg, heat_fx, heat2_fx)
        spm = 1e4 / cfg["microns_per_shell"] / (cfg["mu_a"] + cfg["mu_s"])
        edges = np.arange(cfg["shells"]) / spm
        expect = np.exp(-edges) - np.exp(-(edges + 1.0 / spm))
        expect[-1] = np.exp(-edges[-1])
        sigma = np.sqrt(expect * (1 - expect) / n)
        assert np.abs(heat / n - expect).max() < 5 * sigma.max() + 1e-4
        assert abs(heat.sum() / n - 1.0) < 1e-6


def test_radial_cross_check_mode(gpu, orc):
    """"walk_mode" = 1: the reduced radial walk equals its own CPU replay in every
    integer, and agrees statistically with the 3-D walk (the product) at 2^28 photons per side."""
    gpu.set_option("walk_mode", 1)
    try:
        for name, n in (("default", 20000), ("finegrid", 20000)):
            h, h2 = gpu.photons_fx(name, SEED, 77, n)
            ev = gpu.last_run_info().events
            r, r2, rev = orc.replay(name, SEED, 77, n, mode=1)
            assert ev == rev and int(h.sum()) == int(r.sum()) and int(h2.sum()) == int(r2.sum())
            assert np.abs(h.astype(np.int64) - r.astype(np.int64)).sum() / 2 <= 6e-3 * r.sum()
        radial = gpu_batches(gpu, "default", 32, 1 << 23, seed=99)
    finally:
        gpu.set_option("walk_mode", 0)
    full = gpu_batches(gpu, "default", 32, 1 << 23, seed=99)
    z, ok = batch_means_z(radial[0], 1 << 23, full[0], 1 << 23)
    assert ok.all() and np.abs(z).max() < 4.5 and abs(z.mean()) < 0.6, z
    gpu.set_option("walk_mode", 1)
    gpu.set_option("block_threads", 256)
    try:
        with pytest.raises(gpu.TinyMcError):          # compiled for the default launch shape only
            gpu.photons_fx("default", SEED, 0, 1000)
    finally:
        gpu.set_option("block_threads", 0)
        gpu.set_option("walk_mode", 0)


# ------------------------------------------------------------------ 2. bit-reproducibility
@pytest.mark.parametrize("name,n", [("default", 300000), ("highalbedo", 3000), ("finegrid", 200000)])
def test_result_is_independent_of_split_and_launch_shape(gpu, name, n):
    base = gpu.photons_fx(name, SEED, 0, n)
    # any split of the photon range
    acc = [np.zeros_like(base[0]), np.zeros_like(base[1])]
    for lo, cnt in ((0, n // 7), (n // 7, 1), (n // 7 + 1, n - n // 7 - 1)):
        gpu.photons_fx(name, SEED, lo, cnt, acc[0], acc[1])
    assert np.array_equal(acc[0], base[0]) and np.array_equal(acc[1], base[1])
    # any block shape / residency / drain interval
    shapes = [dict(block_threads=128), dict(block_threads=512, blocks_per_sm=1), dict(flush_iters=5),
              dict(block_threads=1024, flush_iters=17),
              dict(tally_layout=1),    # one histogram per block (integer clamp) instead of per-lane copies (.sat clamp)
              dict(tally_layout=3)]    # one histogram per block, .sat clamp: ONE overflow word takes 20-67 % of the events
    if name == "finegrid":             # auto = layout 3 here (no photon leaves a 180-mean-free-path grid)
        shapes = [dict(block_threads=512), dict(flush_iters=64), dict(block_threads=256, flush_iters=100), dict(tally_layout=1), dict(tally_layout=3)]
    for opts in shapes:
        for k, v in opts.items():
            gpu.set_option(k, v)
        try:
            again = gpu.photons_fx(name, SEED, 0, n)
        finally:
            for k in opts:
                gpu.set_option(k, 0)
        assert np.array_equal(again[0], base[0]) and np.array_equal(again[1], base[1]), opts


def test_run_to_run_identical_and_seed_sensitive(gpu):
    a = gpu.photons_fx("default", 1, 0, 100000)
    b = gpu.photons_fx("default", 1, 0, 100000)
    c = gpu.photons_fx("default", 2, 0, 100000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], c[0])


def test_multi_gpu_is_bit_identical_to_one_gpu(gpu):
    """The library's own N-GPU path (tmc_init(N): shards, in-library ncclReduce or host-side sum) gives the
    words of the one-GPU run, for the plain and for the batched call.  Needs a multi-GPU box
    (gpurun --gpus N); on one GPU the same identity is visible to the driver as bench.py's
    checks.tally_hash, printed at every N."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible; bench.py prints checks.tally_hash at every N for the same identity")
    n = 1 << 22
    one = gpu.photons_fx("default", SEED, 7, n)
    one_b = gpu.photons_fx_batches("finegrid", SEED, 7, n // 4 + 3, 5)
    for g in sorted({2, torch.cuda.device_count()}):
        for nccl in (1, 0):
            gpu.set_option("nccl_reduce", nccl)
            gpu.init(g)
            try:
                many = gpu.photons_fx("default", SEED, 7, n)
                assert gpu.last_run_info().n_gpus == g
                many_b = gpu.photons_fx_batches("finegrid", SEED, 7, n // 4 + 3, 5)
            finally:
                gpu.set_option("nccl_reduce", 1)
                gpu.init(1)
            assert np.array_equal(many[0], one[0]) and np.array_equal(many[1], one[1]), (g, nccl)
            assert np.array_equal(many_b[0], one_b[0]) and np.array_equal(many_b[1], one_b[1]), (g, nccl)


@pytest.mark.parametrize("name,n,nb", [("default", 100003, 7), ("finegrid", 50000, 64), ("highalbedo", 700, 3)])
def test_batches_in_one_pass_equal_separate_calls(gpu, name, n, nb):
    """tmc_photons_fx_batches: batch b is exactly the sub-range a separate tmc_photons_fx call would walk
    (the remainder goes to the first batches), so the batches add up to the unsplit run bit for bit."""
    bh, bh2 = gpu.photons_fx_batches(name, SEED, 11, n, nb)
    info = gpu.last_run_info()
    assert info.photons == n and info.gpu_launches == nb and info.retries == 0
    lo = 11
    for b in range(nb):
        cnt = n // nb + (1 if b < n % nb else 0)
        h, h2 = gpu.photons_fx(name, SEED, lo, cnt)
        assert np.array_equal(h, bh[b]) and np.array_equal(h2, bh2[b]), b
        lo += cnt
    whole = gpu.photons_fx(name, SEED, 11, n)
    assert np.array_equal(bh.sum(axis=0), whole[0]) and np.array_equal(bh2.sum(axis=0), whole[1])
    with pytest.raises(gpu.TinyMcError):
        gpu.photons_fx_batches(name, SEED, 0, 100, 0)


# ------------------------------------------------------------------ 3. statistics vs the reference
def gpu_batches(gpu, name, nb, n, seed=SEED):
    heat, heat2 = [], []
    for b in range(nb):
        hfx, h2fx = gpu.photons_fx(name, seed, b * n, n)
        h, h2 = gpu.capi.fx_to_float64(name, hfx, h2fx)
        heat.append(h)
        heat2.append(h2)
    return np.stack(heat), np.stack(heat2)


# sum heat2 / N against (1-a)/(1+a): the rescaled integer squares are exact to ~1e-5 (default) resp. the
# statistical spread of 2^21 long-lived photons (high albedo)
HEAT2_TOL = {"default": 5e-5, "finegrid": 5e-5, "highalbedo": 5e-4}


def group(a, name):
    return a.reshape(a.shape[0], 128, 128).sum(axis=2) if name == "finegrid" else a


@pytest.mark.parametrize("name,nb,n", [("default", 64, 1 << 22), ("highalbedo", 64, 1 << 15), ("finegrid", 64, 1 << 22)])
def test_every_shell_within_4_sigma_of_the_reference_walk(gpu, name, nb, n):
    """The north-star tolerance: per-shell mean heat within 4 sigma, total absorbed weight to
    1e-4 relative.  Reference side: photon_port.c (bit-identical to reference photon.c) on
    xoshiro256**, 64 batches of 2^21 photons (2^14 for high albedo): 1.3e8 reference photons against
    2.7e8 GPU photons resolve ~0.05 % per shell; sigma: batch means on both sides."""
    ref = np.load(GOLDEN / f"port_xoshiro_batches_{name}.npz")
    n_ref = int(ref["photons_per_batch"])
    heat, heat2 = gpu_batches(gpu, name, nb, n)
    z, ok = batch_means_z(group(heat, name), n, ref["heat"], n_ref, min_mean=1e-4)
    assert ok.sum() >= (101 if name != "finegrid" else 16)
    assert np.abs(z[ok]).max() < 4.0, (np.abs(z).argmax(), z)
    assert abs(z[ok].mean()) < 0.6
    tot_gpu = heat.sum() / (nb * n)
    tot_ref = ref["heat"].sum() / (ref["heat"].shape[0] * n_ref)
    assert abs(tot_gpu - tot_ref) / tot_ref < 1e-4
    assert abs(tot_gpu - 1.0) < 1e-4                                            # roulette is unbiased
    a = albedo(gpu.CONFIGS[name])
    assert abs(heat2.sum() / (nb * n) / ((1 - a) / (1 + a)) - 1.0) < HEAT2_TOL[name]   # sum of squared deposits
    assert abs(heat2.sum() / (nb * n) - ref["heat2"].sum() / (ref["heat2"].shape[0] * n_ref)) / (ref["heat2"].sum() / (ref["heat2"].shape[0] * n_ref)) < 2e-3


@pytest.mark.parametrize("fixture", ["ref_pcg", "port_xoshiro"])
def test_config5_every_5um_shell_against_1e9_reference_photons(gpu, fixture):
    """Config 5 per 5 um shell an order of magnitude deeper: 4.3e9 GPU photons (256 batches of 2^24, one batched
    call) against 1.07e9 reference photons; per-shell standard error ~0.03 % where the shells are populated."""
    ref = np.load(GOLDEN / f"{fixture}_pershell_finegrid_1e9.npz")
    nb, n = 256, 1 << 24
    bh, _ = gpu.photons_fx_batches("finegrid", 0xF1E9, 0, nb * n, nb)
    s1 = 2.0 ** -int(gpu.fx_scales("finegrid").heat_shift)
    per = bh.astype(np.float64) * (s1 / n)
    mean, var = per.mean(axis=0), per.var(axis=0, ddof=1) / nb
    ok = np.maximum(mean, ref["mean"]) >= 1e-5
    assert ok.sum() > 1400
    z = (mean - ref["mean"])[ok] / np.sqrt(var + ref["var_of_mean"])[ok]
    assert np.abs(z).max() < 4.0, (np.abs(z).argmax(), np.abs(z).max())
    assert abs(np.sqrt((z ** 2).mean()) - 1.0) < 0.1 and abs(z.mean()) < 0.25
    assert (np.abs(z) > 3.0).sum() <= 12
    trend = z[: len(z) // 100 * 100].reshape(-1, 100).mean(axis=1)
    assert np.abs(trend).max() < 0.5, trend


@pytest.mark.parametrize("fixture", ["ref_pcg", "port_xoshiro"])
def test_config5_every_5um_shell_within_4_sigma(gpu, fixture):
    """Config 5 at its NATIVE resolution (SHELLS=16384, 5 um: 90.9 shells per mean free path — where a
    23-bit step, 8-bit polar and 8-bit azimuth stream would show first): every shell carrying at least
    1e-5 of a photon's weight (~1500 shells) within 4 sigma of the reference walk, against the UNMODIFIED
    reference on PCG32 and against the port on xoshiro256** (1.3e8 photons each, 256 batches; 2.7e8 GPU
    photons in 256 batches; batch-means sigma on both sides).  Beyond the maximum: z is standard normal
    (rms, |z| > 3 count) and has no trend over radius (means of 100 consecutive shells)."""
    ref = np.load(GOLDEN / f"{fixture}_pershell_finegrid.npz")
    nb, n = 256, 1 << 20
    heat, _ = gpu_batches(gpu, "finegrid", nb, n, seed=0xF19E)
    per = heat / n
    mean, var = per.mean(axis=0), per.var(axis=0, ddof=1) / nb
    ok = np.maximum(mean, ref["mean"]) >= 1e-5
    assert ok.sum() > 1400
    z = (mean - ref["mean"])[ok] / np.sqrt(var + ref["var_of_mean"])[ok]
    assert np.abs(z).max() < 4.0, (np.abs(z).argmax(), np.abs(z).max())
    assert abs(np.sqrt((z ** 2).mean()) - 1.0) < 0.1 and abs(z.mean()) < 0.25
    assert (np.abs(z) > 3.0).sum() <= 12                       # 3.9 expected of ~1500
    trend = z[: len(z) // 100 * 100].reshape(-1, 100).mean(axis=1)
    assert np.abs(trend).max() < 0.5, trend                    # sigma of such a mean is 0.1 (more with shell-to-shell correlation)


def test_sound_generator_fixtures_agree_with_the_gpu_and_libc_does_not(gpu):
    """Default optics against BOTH sound-generator references (unmodified photon.c on PCG32, port on
    xoshiro256**): 4 sigma in every shell; the same GPU tallies against the unmodified reference on libc
    rand() at the same 4.2e6-photon scale fail it (see test_two_sound_generators_agree_where_libc_rand_does_not)."""
    heat, _ = gpu_batches(gpu, "default", 64, 1 << 22, seed=0xBEEF)
    for fixture in ("ref_pcg_batches_default.npz", "port_xoshiro_batches_default.npz"):
        ref = np.load(GOLDEN / fixture)
        z, ok = batch_means_z(heat, 1 << 22, ref["heat"], int(ref["photons_per_batch"]))
        assert ok.all() and np.abs(z).max() < 4.0 and abs(z.mean()) < 0.6, (fixture, z)
    libc = np.load(GOLDEN / "ref_batches_default.npz")
    z, _ = batch_means_z(heat, 1 << 22, libc["heat"], int(libc["photons_per_batch"]))
    assert np.abs(z).max() > 4.0


@pytest.mark.parametrize("fixture", ["ref_pcg", "port_xoshiro"])
def test_default_optics_against_1e9_reference_photons(gpu, fixture):
    """The same 4-sigma bar an order of magnitude deeper: 4.3e9 GPU photons (64 batches of 2^26, one batched call)
    against 1.07e9 photons of the unmodified reference on PCG32 / of the port on xoshiro256**: the per-shell
    standard error is 0.01 % of the shell's heat, the level at which 8-bit direction tables, the 23-bit step or the
    shared step / direction bits of tmc-stream-4 would have to show if they mattered."""
    ref = np.load(GOLDEN / f"{fixture}_batches_default_1e9.npz")
    nb, n = 64, 1 << 26
    bh, bh2 = gpu.photons_fx_batches("default", 0xD1CE, 0, nb * n, nb)
    heat = np.stack([gpu.capi.fx_to_float64("default", bh[b], bh2[b])[0] for b in range(nb)])
    z, ok = batch_means_z(heat, n, ref["heat"], int(ref["photons_per_batch"]))
    assert ok.all()
    assert np.abs(z).max() < 4.0, (np.abs(z).argmax(), z)
    assert abs(z.mean()) < 0.6 and np.sqrt((z ** 2).mean()) < 1.4
    tot_gpu, tot_ref = heat.sum() / (nb * n), ref["heat"].sum() / (ref["heat"].shape[0] * int(ref["photons_per_batch"]))
    assert abs(tot_gpu - tot_ref) < 2e-5


@pytest.mark.parametrize("fixture", ["ref_pcg", "port_xoshiro"])
def test_high_albedo_against_1e7_reference_photons(gpu, fixture):
    """Config 4's optics (7168 events per photon, 67 % of them beyond the grid) 16 times deeper than the 1e6-photon
    fixtures: 6.7e7 GPU photons (64 batches of 2^20 = 4.8e11 events, one batched call) against 1.68e7 photons of
    the unmodified reference on PCG32 / of the port on xoshiro256** (1.2e11 events each): every shell within
    4 sigma at a per-shell standard error of ~0.03 %, total absorbed weight to 1e-4 (north star)."""
    ref = np.load(GOLDEN / f"{fixture}_batches_highalbedo_1e7.npz")
    nb, n = 64, 1 << 20
    bh, bh2 = gpu.photons_fx_batches("highalbedo", 0xA1BED0, 0, nb * n, nb)
    heat = np.stack([gpu.capi.fx_to_float64("highalbedo", bh[b], bh2[b])[0] for b in range(nb)])
    z, ok = batch_means_z(heat, n, ref["heat"], int(ref["photons_per_batch"]))
    assert ok.all()
    assert np.abs(z).max() < 4.0, (np.abs(z).argmax(), z)
    assert abs(z.mean()) < 0.6 and np.sqrt((z ** 2).mean()) < 1.4
    tot_gpu, tot_ref = heat.sum() / (nb * n), ref["heat"].sum() / (ref["heat"].shape[0] * int(ref["photons_per_batch"]))
    assert abs(tot_gpu - tot_ref) < 1e-4 * tot_ref


def test_literal_contract_against_the_unmodified_reference(gpu):
    """Against the UNMODIFIED reference object code on libc rand() at a scale comparable with
    the one it ships with (PHOTONS = 32768, reference params.h:10): one 65536-photon reference
    batch, every shell within 4 sigma — both with batch-means sigma and with the literal
    sigma derived from heat2 (reference tiny_mc.c:64) scaled by its known under-estimate (H5)."""
    libc = np.load(GOLDEN / "ref_batches_default.npz")
    n_ref = int(libc["photons_per_batch"])
    heat, heat2 = gpu_batches(gpu, "default", 32, 1 << 16, seed=77)
    z, ok = batch_means_z(heat, 1 << 16, libc["heat"], n_ref, b_use=1)
    assert ok.all() and np.abs(z).max() < 4.0, z
    zl = literal_sigma_z(heat.sum(axis=0), heat2.sum(axis=0), 32 << 16, libc["heat"][0], libc["heat2"][0], n_ref)
    assert np.isnan(zl[-1])                       # the overflow shell has no literal sigma (H5)
    assert np.nanmax(np.abs(zl[:-1])) < 4.0 * 1.75


def test_libc_rand_bias_is_visible_from_the_gpu_too(gpu):
    """Cross-check of the finding in test_oracle_pinned: at 4.2e6 reference photons the GPU
    (Philox) disagrees with the libc-rand() reference in the same systematic way the xoshiro
    port does, while agreeing with the xoshiro port (test above)."""
    libc = np.load(GOLDEN / "ref_batches_default.npz")
    heat, _ = gpu_batches(gpu, "default", 64, 1 << 20, seed=3)
    z, _ = batch_means_z(heat, 1 << 20, libc["heat"], int(libc["photons_per_batch"]))
    assert z[5:40].mean() < -0.5 and z[60:].mean() > 1.5


def test_batch_means_stderr_against_the_per_photon_estimator(gpu, orc):
    """the standard error the product reports (64 batch means, one tmc_photons_fx_batches call)
    against the per-PHOTON second moment of the CPU replay (orc_replay_per_photon, 2^17 photons): the GPU's
    Var(mean) x N reproduces the per-photon variance of every shell, which the reference's per-event Error
    column (tiny_mc.c:64) under-estimates by 10-50 % and cannot give at all for the overflow shell."""
    n, nb = 1 << 22, 64
    bh, bh2 = gpu.photons_fx_batches("default", 91, 0, n, nb)
    per = np.stack([gpu.capi.fx_to_float64("default", bh[b], bh2[b])[0] for b in range(nb)]) / (n // nb)
    var_photon_gpu = per.var(axis=0, ddof=1) / nb * n
    n_cpu = 1 << 17
    heat, sq = orc.replay_per_photon("default", 91, 0, n_cpu)
    var_photon_cpu = sq / n_cpu - (heat / n_cpu) ** 2
    ratio = var_photon_gpu / var_photon_cpu
    assert 0.5 < ratio.min() and ratio.max() < 1.7 and abs(ratio.mean() - 1.0) < 0.06, ratio
    h, h2 = gpu.capi.fx_to_float64("default", bh.sum(axis=0), bh2.sum(axis=0))
    literal = (h2 - h * h / n) / n                              # per photon, reference tiny_mc.c:64 squared x N
    assert literal[-1] < 0 and (var_photon_gpu[:-1] / literal[:-1]).mean() > 1.3


# ------------------------------------------------------------------ full-size properties
def test_config2_full_size_properties(gpu):
    """BASELINE configs[1]: 2^26 photons, default optics, one GPU — size-independent properties."""
    n = 1 << 26
    heat_fx, heat2_fx = gpu.photons_fx("default", SEED, 0, n)
    info = gpu.last_run_info()
    heat, heat2 = gpu.capi.fx_to_float64("default", heat_fx, heat2_fx)
    assert info.photons == n and info.retries == 0
    assert abs(heat.sum() / n - 1.0) < 6 * 0.00301 / np.sqrt(n) + 2e-6           # E[absorbed] = 1
    assert abs(heat2.sum() / n * 21.0 - 1.0) < 5e-5                              # (1-a)/(1+a) = 1/21
    assert abs(info.events / n - 75.665) < 0.01
    ref = np.load(GOLDEN / "port_xoshiro_batches_default.npz")
    extra_ref = ref["heat"][:, -1].sum() / (ref["heat"].shape[0] * int(ref["photons_per_batch"]))
    assert abs(heat[-1] / n - extra_ref) < 2e-4                                  # "extra" (tiny_mc.c:66)
    # checksum of checksums: the two halves of the range add up to the whole, bit for bit
    a = gpu.photons_fx("default", SEED, 0, n // 2)
    gpu.photons_fx("default", SEED, n // 2, n // 2, a[0], a[1])
    assert np.array_equal(a[0], heat_fx) and np.array_equal(a[1], heat2_fx)


def test_config4_full_size_properties(gpu):
    """BASELINE configs[3]: high albedo (MU_A=0.1, MU_S=100), 2^30 photons = 7.7e12 events:
    size-independent properties of the walk (reference photon.c:32,45-49)."""
    n = 1 << 30
    heat_fx, heat2_fx = gpu.photons_fx("highalbedo", SEED, 0, n)
    info = gpu.last_run_info()
    heat, heat2 = gpu.capi.fx_to_float64("highalbedo", heat_fx, heat2_fx)
    a = albedo(gpu.CONFIGS["highalbedo"])
    assert info.photons == n and info.retries == 0
    assert abs(heat.sum() / n - 1.0) < 2e-6                                      # E[absorbed] = 1 (roulette unbiased)
    assert abs(heat2.sum() / n / ((1 - a) / (1 + a)) - 1.0) < 5e-5
    assert abs(info.events / n - 7168.0) < 4.0                                   # 6912 + 2304 / 9 (deterministic schedule)
    ref = np.load(GOLDEN / "port_xoshiro_batches_highalbedo.npz")
    extra_ref = ref["heat"][:, -1].sum() / (ref["heat"].shape[0] * int(ref["photons_per_batch"]))
    assert abs(heat[-1] / n - extra_ref) < 2e-3                                  # overflow-bin share ~0.235


def test_config5_full_size_properties(gpu):
    """BASELINE configs[4]: SHELLS=16384, 5 um shells, 2^30 photons (one shared histogram per block)."""
    n = 1 << 30
    heat_fx, heat2_fx = gpu.photons_fx("finegrid", SEED, 0, n)
    info = gpu.last_run_info()
    heat, heat2 = gpu.capi.fx_to_float64("finegrid", heat_fx, heat2_fx)
    assert info.photons == n and info.retries == 0
    assert abs(heat.sum() / n - 1.0) < 2e-6
    assert abs(heat2.sum() / n * 21.0 - 1.0) < 5e-5
    assert abs(info.events / n - 75.665) < 0.01
    assert heat[-1] / n < 1e-6                                                   # 8.2 cm grid: the overflow bin stays empty
    # the fine grid is the default grid refined 10x: regrouped it is the same profile as config 2
    coarse, _ = gpu.capi.fx_to_float64("default", *gpu.photons_fx("default", SEED + 1, 0, 1 << 26))
    mine = heat[:1000].reshape(100, 10).sum(axis=1) / n
    assert np.abs(mine / (coarse[:100] / (1 << 26)) - 1.0).max() < 5e-3


# ------------------------------------------------------------------ the reference-facing calls
def test_photons_adds_into_caller_arrays_like_photon_does(gpu):
    """tmc_photons() is `for (...) photon(heat, heat2)`: it only ADDS (reference photon.c:30-31)."""
    n = 50000
    heat = np.full(101, 2.0, np.float32)
    heat2 = np.full(101, 1.0, np.float32)
    gpu.photons("default", SEED, 0, n, heat, heat2)
    hfx, h2fx = gpu.photons_fx("default", SEED, 0, n)
    h, h2 = gpu.capi.fx_to_float64("default", hfx, h2fx)
    assert np.array_equal(heat, (np.float32(2.0) + h.astype(np.float32)).astype(np.float32))
    assert np.array_equal(heat2, (np.float32(1.0) + h2.astype(np.float32)).astype(np.float32))


def test_bad_arguments_on_a_gpu(gpu):
    lib = gpu.load()
    p = gpu.capi.make_params("default")
    heat = np.zeros(101, np.float32)
    assert lib.tmc_photons(C.byref(p), 1, 0, 10, None, heat.ctypes.data) == 2
    assert lib.tmc_photons(None, 1, 0, 10, heat.ctypes.data, heat.ctypes.data) == 2
    big = gpu.Params(100000, 2.0, 20.0, 50.0)
    h = np.zeros(100000, np.float32)
    assert lib.tmc_photons(C.byref(big), 1, 0, 10, h.ctypes.data, h.ctypes.data) == 2   # SHELLS beyond shared memory
    assert b"shared memory" in lib.tmc_last_error()


def test_parameter_sweep_evicts_cached_tables_without_changing_results(gpu):
    """The per-device deposit-table cache is bounded (16 optics): a sweep over 40 optics evicts the early
    tables; walking the first optics again rebuilds its table and gives the same words."""
    first = dict(shells=64, mu_a=1.0, mu_s=10.0, microns_per_shell=100.0)
    base = gpu.photons_fx(first, 3, 0, 5000)
    for k in range(40):
        cfg = dict(shells=64, mu_a=1.0 + 0.05 * (k + 1), mu_s=10.0, microns_per_shell=100.0)
        h, _ = gpu.photons_fx(cfg, 3, 0, 500)
        assert gpu.last_run_info().photons == 500 and h.sum() > 0
    again = gpu.photons_fx(first, 3, 0, 5000)
    assert np.array_equal(again[0], base[0]) and np.array_equal(again[1], base[1])


def test_tally_range_tripwire_retries_then_fails_loudly(gpu):
    """A drained u32 word at or above 2^tally_check_bits makes the host repeat the range with an 8x
    shorter drain interval; if that never helps the call fails with TMC_ERR_TALLY_RANGE instead of
    returning tallies that may have wrapped."""
    base = gpu.photons_fx("default", SEED, 0, 1 << 20)
    outcomes = {}
    for bits in range(26, 11, -2):
        gpu.set_option("tally_check_bits", bits)
        try:
            again = gpu.photons_fx("default", SEED, 0, 1 << 20)
            outcomes[bits] = gpu.last_run_info().retries
            assert np.array_equal(again[0], base[0]) and np.array_equal(again[1], base[1]), bits
        except gpu.TinyMcError as e:
            assert e.code == 5
            outcomes[bits] = -e.code
        finally:
            gpu.set_option("tally_check_bits", 0)
    assert outcomes[26] == 0                         # far above what 2^20 photons can pile up
    assert any(v >= 1 for v in outcomes.values()), outcomes      # some threshold is rescued by shorter intervals
    assert outcomes[12] == -5, outcomes              # below a single deposit: fails loudly with TMC_ERR_TALLY_RANGE


def test_device_resident_call_on_a_torch_stream(gpu):
    import torch

    shells, n = 101, 200000
    buf = torch.zeros(2 * shells + 4, dtype=torch.int64, device="cuda:0")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gpu.photons_device("default", SEED, 0, n // 2, 0, buf.data_ptr(), s.cuda_stream)
        gpu.photons_device("default", SEED, n // 2, n // 2, 0, buf.data_ptr(), s.cuda_stream)
    s.synchronize()
    words = buf.cpu().numpy().astype(np.uint64)
    hfx, h2fx = gpu.photons_fx("default", SEED, 0, n)
    assert np.array_equal(words[:shells], hfx) and np.array_equal(words[shells:2 * shells], h2fx)
    assert words[2 * shells] == gpu.last_run_info().events and words[2 * shells + 1] == n and words[2 * shells + 2] == 0
    # the mandatory validity test of the device path: clean here, TMC_ERR_TALLY_RANGE once the tripwire fired
    gpu.device_tallies_check("default", 0, buf.data_ptr(), s.cuda_stream)
    gpu.set_option("tally_check_bits", 12)
    try:
        with torch.cuda.stream(s):
            gpu.photons_device("default", SEED, 0, n, 0, buf.data_ptr(), s.cuda_stream)
        with pytest.raises(gpu.TinyMcError) as err:
            gpu.device_tallies_check("default", 0, buf.data_ptr(), s.cuda_stream)
        assert err.value.code == 5
    finally:
        gpu.set_option("tally_check_bits", 0)


def test_headless_program_prints_the_reference_layout(gpu):
    """The C host program (tiny_mc_b200/host/tiny_mc.c) against the reference's own stdout."""
    exe = ROOT / "tiny_mc_b200" / "bin" / "headless"
    js = ROOT / "gpurun_out" / "headless_test.json"
    js.parent.mkdir(exist_ok=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True,
                         env={**os.environ, "TMC_GPUS": "1", "TMC_JSON": str(js)}).stdout.splitlines()
    import json
    doc = json.loads(js.read_text())
    assert doc["photons"] == 32768 and doc["shells"] == 101 and len(doc["heat"]) == 101
    assert abs(sum(doc["heat"]) / doc["photons"] - 1.0) < 5 * 0.00301 / np.sqrt(32768)
    assert sum(doc["heat_fx"]) == round(sum(doc["heat"]) * 2 ** doc["heat_shift"])
    gold = (GOLDEN / "headless_asshipped.txt").read_text().splitlines()
    assert len(out) == len(gold)
    assert out[:2] == gold[:2] and out[3:7] == gold[3:7] and out[9:11] == gold[9:11]
    assert out[5] == "# Photons    =    32768"
    rows = np.array([[float(v) for v in line.split("\t")] for line in out[11:-1]])
    ref = np.array([[float(v) for v in line.split("\t")] for line in gold[11:-1]])
    assert np.array_equal(rows[:, 0], ref[:, 0])                                  # radii
    # heat column: both are 32768-photon estimates of the same profile; Error column is the 1-sigma
    zz = (rows[:, 1] - ref[:, 1]) / np.sqrt(rows[:, 2] ** 2 + ref[:, 2] ** 2)
    assert np.abs(zz).max() < 4.0 * 1.75 and abs(zz.mean()) < 0.6
    assert out[-1].startswith("# extra\t") and abs(float(out[-1].split("\t")[1]) - 0.0235) < 3e-3


def test_frames_program_accumulates_like_the_viewer(gpu):
    """Incremental use of the tallies (reference cg_mc.c:71-87): 16 frames of 4096 photons into the
    same running arrays give exactly the one-shot result."""
    exe = ROOT / "tiny_mc_b200" / "bin" / "frames"
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    frames = [line for line in out if line.startswith("frame")]
    assert len(frames) == 16 and frames[-1].split("\t")[1] == "photons 65536"
    gpu.init(1)
    hfx, h2fx = gpu.photons_fx("default", 4242, 0, 65536)
    checksum = 0
    for a, b in zip(hfx.tolist(), h2fx.tolist()):
        checksum = (checksum * 1000003 + a + 31 * b) % (1 << 64)
    assert out[-1] == f"# checksum\t{checksum}"
    assert abs(float(frames[-1].split("absorbed/photon ")[1]) - 1.0) < 5 * 0.00301 / np.sqrt(65536)


def test_photon_compat_shim_is_one_photon_per_call(gpu):
    """`void photon(float*, float*)` (reference photon.h:3): k calls == photons [0, k)."""
    lib = C.CDLL(str(ROOT / "tiny_mc_b200" / "lib" / "libphoton_compat.so"))
    lib.photon_seed.argtypes = [C.c_ulonglong]
    lib.photon.argtypes = [C.c_void_p, C.c_void_p]
    heat = np.zeros(101, np.float32)
    heat2 = np.zeros(101, np.float32)
    lib.photon_seed(4242)
    for _ in range(8):
        lib.photon(heat.ctypes.data, heat2.ctypes.data)
    gpu.init(1)     # the shim re-initialised the library
    hfx, h2fx = gpu.photons_fx("default", 4242, 0, 8)
    h, _ = gpu.capi.fx_to_float64("default", hfx, h2fx)
    assert abs(float(heat.sum()) - h.sum()) < 1e-4 and np.allclose(heat, h, atol=1e-5)
