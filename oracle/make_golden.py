#!/usr/bin/env python
"""oracle/make_golden.py — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Regenerates tests/golden/*.json|npz from the UNMODIFIED reference object code in oracle/_ref
(run `make -C oracle ref` first; needs /root/reference, so this runs in the build container
only — the fixtures are what travels to the GPU box and into git).

  ref_float_tallies.json   reference driver semantics (tiny_mc.c:43,47-49): srand(seed), N calls
                           of photon() into ONE pair of float arrays; the float bits, per config.
                           Pins oracle/photon_port.c bit for bit.
  ref_batches_<cfg>.npz    B independent srand seeds x n photons, 256-photon float chunks summed
                           in double: per-batch heat / heat2.  The statistical
                           reference for the GPU parity tests (batch-means sigma).
  port_xoshiro_batches_<cfg>.npz  the same batches from oracle/photon_port.c (pinned to the
                           reference bit for bit) driven by xoshiro256** instead of glibc rand():
                           the high-statistics reference, free of rand()'s lag-3/31 correlation.
  ref_pcg_batches_<cfg>.npz  the same batches from the UNMODIFIED photon.c compiled with -Drand=pcg31
                           (oracle/pcg31.c: PCG32 bound to the reference's rand() calls at compile time):
                           a second sound generator under the reference's own object code.
  *_pershell_finegrid.npz  config 5 at its native 5 um resolution: per-shell mean and batch-means
                           variance of the mean (256 batches of 2^19 photons), all 16384 shells,
                           from the xoshiro port and from the unmodified reference on PCG32.
  *_batches_default_1e9.npz  default optics at 64 x 2^24 = 1.07e9 photons per reference (unmodified reference on
                           PCG32, port on xoshiro256**): per-shell precision 0.01 %.  Opt-in: `make_golden.py default_1e9`.
  *_batches_highalbedo_1e7.npz  config 4's optics at 64 x 2^18 = 1.68e7 photons per reference (7168 events per photon:
                           1.2e11 events each).  Opt-in: `make_golden.py highalbedo_1e7` (30 minutes on 8 cores).
  *_pershell_finegrid_1e9.npz  config 5 per 5 um shell at 256 x 2^22 = 1.07e9 photons per reference.  Opt-in:
                           `make_golden.py finegrid_pershell_1e9` (20 minutes on 8 cores).
  headless_asshipped.txt   stdout of the reference `headless` built exactly as its Makefile does,
                           SEED=20141017; with ref_float_tallies.json["headless"] (same seed, same
                           32768 photons) it pins the printout formatter byte for byte.
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import pyoracle as orc  # noqa: E402

GOLD = Path(__file__).resolve().parent.parent / "tests" / "golden"
HEADLESS_SEED = 20141017


def bits(a):
    return [int(v) for v in np.asarray(a, np.float32).view(np.uint32)]


def main():
    assert orc.have_ref(), "run `make -C oracle ref` first"
    GOLD.mkdir(parents=True, exist_ok=True)

    only = set(sys.argv[1:])
    exact = {}
    for name, seed, n in (("default", 1234, 4096), ("highalbedo", 99, 64), ("finegrid", 4321, 4096)):
        r = orc.run_batch(name, seed, n, chunk=0, impl="reference")
        nz = np.nonzero(r["heat_f"])[0]
        exact[name] = dict(config=orc.CONFIGS[name], seed=seed, photons=n,
                           nonzero_shells=[int(i) for i in nz],
                           heat_bits=bits(r["heat_f"][nz]), heat2_bits=bits(r["heat2_f"][nz]))
    r = orc.run_batch("default", HEADLESS_SEED, 32768, chunk=0, impl="reference")
    exact["headless"] = dict(config=orc.CONFIGS["default"], seed=HEADLESS_SEED, photons=32768,
                             heat_bits=bits(r["heat_f"]), heat2_bits=bits(r["heat2_f"]))
    # the unmodified reference with rand() bound to PCG32: pins pcg31.c and the port's third generator
    r = orc.run_batch("default", 2718, 4096, chunk=0, impl="reference_pcg")
    exact["default_pcg"] = dict(config=orc.CONFIGS["default"], seed=2718, photons=4096, rng="pcg",
                                heat_bits=bits(r["heat_f"]), heat2_bits=bits(r["heat2_f"]))
    (GOLD / "ref_float_tallies.json").write_text(json.dumps(exact))

    out = subprocess.run([str(orc.REF_DIR / "headless_asshipped")], capture_output=True, text=True, check=True).stdout
    (GOLD / "headless_asshipped.txt").write_text(out)

    # chunk = photons per fresh float tally.  High albedo makes ~7150 deposits of
    # ~1e-3 * w per photon: even 256-photon float chunks lose 1.3e-4 of the weight, so 8 there.
    # n = photons per batch of the libc-rand() reference, n_good = of the xoshiro port (32x / 16x more:
    # 1.3e8 photons resolve ~0.05 % per shell, the level at which the product's 42-bit events could show;
    # a 4.2e6-photon sample once sat 3.5 sigma low over shells 60-76 and failed a correct kernel).
    plans = (("default", 64, 1 << 16, 1 << 21, 256), ("highalbedo", 64, 1 << 10, 1 << 14, 8),
             ("finegrid", 64, 1 << 16, 1 << 21, 256))
    only = set(sys.argv[1:])
    for name, nb, n, n_good, chunk in plans:
        if only and name not in only:
            continue
        seeds = [1000 + 7 * b for b in range(nb)]
        heat, heat2, _, secs, wall = orc.run_batches(name, seeds, n, chunk=chunk, impl="reference")
        if name == "finegrid":   # 16384 shells x 64 batches is too big to commit: keep 128-shell groups
            heat = heat.reshape(nb, 128, 128).sum(axis=2)
            heat2 = heat2.reshape(nb, 128, 128).sum(axis=2)
        np.savez_compressed(GOLD / f"ref_batches_{name}.npz", heat=heat, heat2=heat2, seeds=np.array(seeds),
                            photons_per_batch=n, chunk=chunk)
        print(f"{name}: {nb} x {n} photons, cpu {secs.sum():.1f} s, wall {wall:.1f} s, "
              f"total/photon {heat.sum() / (nb * n):.6f}")
        # the same walk code (photon_port.c, pinned to the reference above) on a sound generator
        n = n_good
        heat, heat2, _, secs, wall = orc.run_batches(name, seeds, n, chunk=chunk, impl="port", rng="xoshiro")
        if name == "finegrid":
            heat = heat.reshape(nb, 128, 128).sum(axis=2)
            heat2 = heat2.reshape(nb, 128, 128).sum(axis=2)
        np.savez_compressed(GOLD / f"port_xoshiro_batches_{name}.npz", heat=heat, heat2=heat2, seeds=np.array(seeds),
                            photons_per_batch=n, chunk=chunk)
        print(f"{name} (xoshiro): total/photon {heat.sum() / (nb * n):.6f}")
        # the UNMODIFIED reference object code on a second sound generator (rand() -> PCG32 at compile time)
        heat, heat2, _, secs, wall = orc.run_batches(name, seeds, n, chunk=chunk, impl="reference_pcg")
        if name == "finegrid":
            heat = heat.reshape(nb, 128, 128).sum(axis=2)
            heat2 = heat2.reshape(nb, 128, 128).sum(axis=2)
        np.savez_compressed(GOLD / f"ref_pcg_batches_{name}.npz", heat=heat, heat2=heat2, seeds=np.array(seeds),
                            photons_per_batch=n, chunk=chunk)
        print(f"{name} (unmodified reference on pcg32): total/photon {heat.sum() / (nb * n):.6f}, wall {wall:.1f} s")
    if "default_1e9" in only:      # opt-in (25 minutes on 8 cores): default optics at 1.07e9 photons per reference
        nb, n = 64, 1 << 24
        seeds = [70000 + 13 * b for b in range(nb)]
        for tag, kw in (("ref_pcg", dict(impl="reference_pcg")), ("port_xoshiro", dict(impl="port", rng="xoshiro"))):
            heat, heat2, _, secs, wall = orc.run_batches("default", seeds, n, chunk=256, **kw)
            np.savez_compressed(GOLD / f"{tag}_batches_default_1e9.npz", heat=heat, heat2=heat2, seeds=np.array(seeds),
                                photons_per_batch=n, chunk=256)
            print(f"default 1e9 ({tag}): total/photon {heat.sum() / (nb * n):.7f}, wall {wall:.1f} s")
    if "highalbedo_1e7" in only:   # opt-in (30 minutes on 8 cores): config 4's optics at 64 x 2^18 = 1.68e7 photons per reference
        nb, n = 64, 1 << 18
        seeds = [110000 + 19 * b for b in range(nb)]
        for tag, kw in (("ref_pcg", dict(impl="reference_pcg")), ("port_xoshiro", dict(impl="port", rng="xoshiro"))):
            heat, heat2, _, secs, wall = orc.run_batches("highalbedo", seeds, n, chunk=8, **kw)
            np.savez_compressed(GOLD / f"{tag}_batches_highalbedo_1e7.npz", heat=heat, heat2=heat2, seeds=np.array(seeds),
                                photons_per_batch=n, chunk=8)
            print(f"highalbedo 1e7 ({tag}): total/photon {heat.sum() / (nb * n):.7f}, wall {wall:.1f} s", flush=True)
    if "finegrid_pershell_1e9" in only:      # opt-in (20 minutes on 8 cores): config 5 per 5 um shell at 1.07e9 photons per reference
        nb, n = 256, 1 << 22
        seeds = [90000 + 17 * b for b in range(nb)]
        for tag, kw in (("ref_pcg", dict(impl="reference_pcg")), ("port_xoshiro", dict(impl="port", rng="xoshiro"))):
            heat, heat2, _, secs, wall = orc.run_batches("finegrid", seeds, n, chunk=256, **kw)
            per = heat / n
            np.savez_compressed(GOLD / f"{tag}_pershell_finegrid_1e9.npz", mean=per.mean(axis=0),
                                var_of_mean=per.var(axis=0, ddof=1) / nb, heat2_mean=(heat2 / n).mean(axis=0),
                                batches=nb, photons_per_batch=n, seeds=np.array(seeds))
            print(f"finegrid per shell 1e9 ({tag}): total/photon {per.sum(axis=1).mean():.7f}, wall {wall:.1f} s")
    if not only or "finegrid_pershell" in only:
        # config 5 per 5 um shell: mean and variance of the mean from 256 batches of 2^19 photons
        nb, n = 256, 1 << 19
        seeds = [5000 + 11 * b for b in range(nb)]
        for tag, kw in (("port_xoshiro", dict(impl="port", rng="xoshiro")), ("ref_pcg", dict(impl="reference_pcg"))):
            heat, heat2, _, secs, wall = orc.run_batches("finegrid", seeds, n, chunk=256, **kw)
            per = heat / n
            np.savez_compressed(GOLD / f"{tag}_pershell_finegrid.npz", mean=per.mean(axis=0),
                                var_of_mean=per.var(axis=0, ddof=1) / nb, heat2_mean=(heat2 / n).mean(axis=0),
                                batches=nb, photons_per_batch=n, seeds=np.array(seeds))
            print(f"finegrid per shell ({tag}): total/photon {per.sum(axis=1).mean():.6f}, wall {wall:.1f} s")


if __name__ == "__main__":
    main()
