#!/usr/bin/env python
"""tools/parity_report.py — the numbers behind tests/test_gpu_parity.py::test_every_shell_within_4_sigma...:
per configuration max |z|, mean z, shells beyond 3 sigma, total absorbed weight of both sides (run on a B200)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import tiny_mc_b200 as tmc  # noqa: E402
from stats import batch_means_z, literal_sigma_z  # noqa: E402

tmc.init(1)
out = {}
GOLD = ROOT / "tests" / "golden"
# config 5 per 5 um shell, against both sound-generator references (256 batches of 2^20 GPU photons)
nb, n = 256, 1 << 20
bh, bh2 = tmc.photons_fx_batches("finegrid", 0xF19E, 0, nb * n, nb)
per = np.stack([tmc.capi.fx_to_float64("finegrid", bh[b], bh2[b])[0] for b in range(nb)]) / n
mean, var = per.mean(axis=0), per.var(axis=0, ddof=1) / nb
for fixture in ("ref_pcg", "port_xoshiro"):
    ref = np.load(GOLD / f"{fixture}_pershell_finegrid.npz")
    ok = np.maximum(mean, ref["mean"]) >= 1e-5
    z = (mean - ref["mean"])[ok] / np.sqrt(var + ref["var_of_mean"])[ok]
    trend = z[: len(z) // 100 * 100].reshape(-1, 100).mean(axis=1)
    out[f"finegrid_per_shell_vs_{fixture}"] = dict(
        gpu_photons=nb * n, reference_photons=int(ref["batches"]) * int(ref["photons_per_batch"]), shells_tested=int(ok.sum()),
        max_abs_z=float(np.abs(z).max()), rms_z=float(np.sqrt((z ** 2).mean())), mean_z=float(z.mean()),
        shells_beyond_3_sigma=int((np.abs(z) > 3).sum()), mean_z_per_100_shells=[round(float(v), 2) for v in trend],
        median_relative_sigma_per_shell=float(np.median(np.sqrt(var + ref["var_of_mean"])[ok] / mean[ok])))
# config 5 per 5 um shell against the 1.07e9-photon references (4.3e9 GPU photons in 256 batches)
nb, n = 256, 1 << 24
bh, _ = tmc.photons_fx_batches("finegrid", 0xF1E9, 0, nb * n, nb)
per = bh.astype(np.float64) * (2.0 ** -int(tmc.fx_scales("finegrid").heat_shift) / n)
mean, var = per.mean(axis=0), per.var(axis=0, ddof=1) / nb
for fixture in ("ref_pcg", "port_xoshiro"):
    ref = np.load(GOLD / f"{fixture}_pershell_finegrid_1e9.npz")
    ok = np.maximum(mean, ref["mean"]) >= 1e-5
    z = (mean - ref["mean"])[ok] / np.sqrt(var + ref["var_of_mean"])[ok]
    trend = z[: len(z) // 100 * 100].reshape(-1, 100).mean(axis=1)
    out[f"finegrid_per_shell_1e9_vs_{fixture}"] = dict(
        gpu_photons=nb * n, reference_photons=int(ref["batches"]) * int(ref["photons_per_batch"]), shells_tested=int(ok.sum()),
        max_abs_z=float(np.abs(z).max()), rms_z=float(np.sqrt((z ** 2).mean())), mean_z=float(z.mean()),
        shells_beyond_3_sigma=int((np.abs(z) > 3).sum()), mean_z_per_100_shells=[round(float(v), 2) for v in trend],
        median_relative_sigma_per_shell=float(np.median(np.sqrt(var + ref["var_of_mean"])[ok] / mean[ok])))
# default optics against the 1.07e9-photon references (4.3e9 GPU photons)
nb, n = 64, 1 << 26
bh, bh2 = tmc.photons_fx_batches("default", 0xD1CE, 0, nb * n, nb)
heat9 = np.stack([tmc.capi.fx_to_float64("default", bh[b], bh2[b])[0] for b in range(nb)])
for fixture in ("ref_pcg", "port_xoshiro"):
    ref = np.load(GOLD / f"{fixture}_batches_default_1e9.npz")
    n_ref = int(ref["photons_per_batch"])
    z, ok = batch_means_z(heat9, n, ref["heat"], n_ref)
    rel = np.sqrt((heat9 / n).var(axis=0, ddof=1) / nb + (ref["heat"] / n_ref).var(axis=0, ddof=1) / ref["heat"].shape[0]) / (heat9 / n).mean(axis=0)
    out[f"default_1e9_vs_{fixture}"] = dict(gpu_photons=nb * n, reference_photons=ref["heat"].shape[0] * n_ref, shells_tested=int(ok.sum()),
                                            max_abs_z=float(np.abs(z).max()), rms_z=float(np.sqrt((z ** 2).mean())), mean_z=float(z.mean()),
                                            shells_beyond_3_sigma=int((np.abs(z) > 3).sum()), median_relative_sigma_per_shell=float(np.median(rel)),
                                            absorbed_per_photon_gpu=float(heat9.sum() / (nb * n)),
                                            absorbed_per_photon_reference=float(ref["heat"].sum() / (ref["heat"].shape[0] * n_ref)),
                                            z=[round(float(v), 2) for v in z])
# config 4's optics against the 1.68e7-photon references (6.7e7 GPU photons = 4.8e11 events)
if (GOLD / "ref_pcg_batches_highalbedo_1e7.npz").exists():
    nb, n = 64, 1 << 20
    bh, bh2 = tmc.photons_fx_batches("highalbedo", 0xA1BED0, 0, nb * n, nb)
    heat7 = np.stack([tmc.capi.fx_to_float64("highalbedo", bh[b], bh2[b])[0] for b in range(nb)])
    for fixture in ("ref_pcg", "port_xoshiro"):
        ref = np.load(GOLD / f"{fixture}_batches_highalbedo_1e7.npz")
        n_ref = int(ref["photons_per_batch"])
        z, ok = batch_means_z(heat7, n, ref["heat"], n_ref)
        rel = np.sqrt((heat7 / n).var(axis=0, ddof=1) / nb + (ref["heat"] / n_ref).var(axis=0, ddof=1) / ref["heat"].shape[0]) / (heat7 / n).mean(axis=0)
        out[f"highalbedo_1e7_vs_{fixture}"] = dict(gpu_photons=nb * n, reference_photons=ref["heat"].shape[0] * n_ref, shells_tested=int(ok.sum()),
                                                   max_abs_z=float(np.abs(z).max()), rms_z=float(np.sqrt((z ** 2).mean())), mean_z=float(z.mean()),
                                                   shells_beyond_3_sigma=int((np.abs(z) > 3).sum()), median_relative_sigma_per_shell=float(np.median(rel)),
                                                   absorbed_per_photon_gpu=float(heat7.sum() / (nb * n)),
                                                   absorbed_per_photon_reference=float(ref["heat"].sum() / (ref["heat"].shape[0] * n_ref)),
                                                   z=[round(float(v), 2) for v in z])
for name, nb, n in (("default", 64, 1 << 22), ("highalbedo", 64, 1 << 15), ("finegrid", 64, 1 << 22)):
    ref = np.load(GOLD / f"port_xoshiro_batches_{name}.npz")
    n_ref = int(ref["photons_per_batch"])
    bh, bh2 = tmc.photons_fx_batches(name, 0x5EED, 0, nb * n, nb)
    heat = np.stack([tmc.capi.fx_to_float64(name, bh[b], bh2[b])[0] for b in range(nb)])
    heat2 = np.stack([tmc.capi.fx_to_float64(name, bh[b], bh2[b])[1] for b in range(nb)])
    pcg = np.load(GOLD / f"ref_pcg_batches_{name}.npz")
    gg = heat.reshape(nb, 128, 128).sum(axis=2) if name == "finegrid" else heat
    zp, okp = batch_means_z(gg, n, pcg["heat"], int(pcg["photons_per_batch"]), min_mean=1e-4)
    out[f"{name}_vs_unmodified_reference_on_pcg32"] = dict(shells_tested=int(okp.sum()), max_abs_z=float(np.abs(zp[okp]).max()),
                                                          rms_z=float(np.sqrt((zp[okp] ** 2).mean())), mean_z=float(zp[okp].mean()))
    g = heat.reshape(nb, 128, 128).sum(axis=2) if name == "finegrid" else heat
    z, ok = batch_means_z(g, n, ref["heat"], n_ref, min_mean=1e-4)
    rel_sigma = np.sqrt((g / n).var(axis=0, ddof=1) / nb + (ref["heat"] / n_ref).var(axis=0, ddof=1) / ref["heat"].shape[0])[ok] / (g / n).mean(axis=0)[ok]
    zl = literal_sigma_z(heat.sum(0), heat2.sum(0), nb * n, ref["heat"].sum(0) if name != "finegrid" else heat.sum(0), ref["heat2"].sum(0) if name != "finegrid" else heat2.sum(0), 64 * n_ref if name != "finegrid" else nb * n)
    out[name] = dict(gpu_photons=nb * n, reference_photons=64 * n_ref, shells_tested=int(ok.sum()), max_abs_z=float(np.abs(z[ok]).max()),
                     mean_z=float(z[ok].mean()), shells_beyond_3_sigma=int((np.abs(z[ok]) > 3).sum()),
                     median_relative_sigma_per_shell=float(np.median(rel_sigma)),
                     absorbed_per_photon_gpu=float(heat.sum() / (nb * n)), absorbed_per_photon_reference=float(ref["heat"].sum() / (64 * n_ref)),
                     literal_heat2_sigma_max_abs_z=(None if name == "finegrid" else float(np.nanmax(np.abs(zl[:-1])))),
                     literal_heat2_sigma_nan_shells=(None if name == "finegrid" else int(np.isnan(zl).sum())),
                     z=[round(float(v), 2) for v in z[ok]])
tmc.finalize()
print(json.dumps(out, indent=1))
