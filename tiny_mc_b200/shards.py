"""Photon-range sharding for one-process-per-GPU hosts (torchrun ranks).

The reference has one loop `for (i = 0; i < PHOTONS; ++i) photon(heat, heat2)` (reference
tiny_mc.c:47-49).  Photons are independent, and the product's stream is keyed by the GLOBAL
photon index, so a job of photons [first, first + n) splits into `world` contiguous shards,
each rank walks its own shard into its own tally buffer, and ONE exact integer sum of the
2*SHELLS+4 tally words (NCCL all-reduce on GPUs, gloo in the CPU tests) gives the same words
as a single-rank run.  The library's own multi-GPU path (tmc_photons with tmc_init(n > 1))
uses the same split (tmc_api.cu: run_range).
"""
from __future__ import annotations


def shard_range(first_photon: int, n_photons: int, rank: int, world: int):
    """(first, count) of rank's contiguous shard; the first n % world ranks take one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    if n_photons < 0 or first_photon < 0:
        raise ValueError("negative photon range")
    base, extra = divmod(n_photons, world)
    lo = first_photon + base * rank + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def tally_words(shells: int) -> int:
    """Length of the u64 tally buffer: heat_fx | heat2_fx | events, photons, range flag, reserved."""
    return 2 * shells + 4
