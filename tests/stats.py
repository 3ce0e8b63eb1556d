"""Statistical comparison helpers for the parity tests."""
import numpy as np


def batch_means_z(a_batches, a_n, b_batches, b_n, min_mean=0.0, b_use=None):
    """Per-shell z of the difference of per-photon means, sigma from batch means on both sides.

    a_batches, b_batches: [B, S] tallies (weight units) of batches of a_n / b_n photons each.
    b_use: compare against the mean of only the first `b_use` batches of b (a smaller reference
    sample), while the per-batch variance is still estimated from ALL batches of b so that the
    statistic keeps its degrees of freedom.
    """
    a = np.asarray(a_batches, np.float64) / a_n
    b = np.asarray(b_batches, np.float64) / b_n
    k = b.shape[0] if b_use is None else int(b_use)
    ma, mb = a.mean(axis=0), b[:k].mean(axis=0)
    va = a.var(axis=0, ddof=1) / a.shape[0]
    vb = b.var(axis=0, ddof=1) / k
    sd = np.sqrt(va + vb)
    ok = (sd > 0) & (np.maximum(ma, mb) > min_mean)
    z = np.zeros_like(ma)
    z[ok] = (ma[ok] - mb[ok]) / sd[ok]
    return z, ok


def literal_sigma_z(heat_a, heat2_a, n_a, heat_b, heat2_b, n_b):
    """The contract's literal statistic: sigma derived from heat2 as reference tiny_mc.c:64 does,
    sigma_X^2 = (heat2_X - heat_X^2 / N_X) / N_X^2.  NaN where that variance is negative (H5)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        va = (heat2_a - heat_a**2 / n_a) / n_a**2
        vb = (heat2_b - heat_b**2 / n_b) / n_b**2
        z = (heat_a / n_a - heat_b / n_b) / np.sqrt(va + vb)
    return z


def fnv64(words) -> str:
    """FNV-1a (64 bit) over the little-endian bytes of u64 tally words (the same hash as bench.py's checks.tally_hash)."""
    h = 0xCBF29CE484222325
    for b in np.asarray(words).astype("<u8").tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"
