"""tiny_mc_b200 — B200-native photon random walk behind tiny_mc's C API.

The product is the C-ABI shared library ``tiny_mc_b200/lib/libtinymc_b200.so``
(``include/tiny_mc_b200.h``) and the C host program ``tiny_mc_b200/host/tiny_mc.c``.
This Python package is only the ctypes binding used by the tests and ``bench.py``;
it never computes anything itself and raises if the CUDA library is missing.
"""
from .capi import (  # noqa: F401
    CONFIGS,
    Params,
    RunInfo,
    Scales,
    TinyMcError,
    device_tallies_check,
    fx_accumulate,
    fx_scales,
    generation_plan,
    init,
    finalize,
    last_run_info,
    lib_path,
    load,
    photons,
    photons_device,
    photons_fx,
    photons_fx_batches,
    prepare,
    set_option,
)
