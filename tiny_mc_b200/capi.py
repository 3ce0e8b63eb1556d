"""ctypes binding of include/tiny_mc_b200.h (the drop-in boundary for reference photon.h:3).

Mirrors the C entry points one to one; names follow the reference's domain
(photons, heats, heats_squared, shells).  No computation happens here and there is
no fallback: if the CUDA library cannot be loaded, or no B200 is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parent


class TinyMcError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"tiny_mc_b200 error {code}: {message}")
        self.code = code


class Params(C.Structure):
    """Run-time form of reference params.h:5-23 (tmc_params)."""

    _fields_ = [
        ("shells", C.c_uint32),
        ("mu_a", C.c_float),
        ("mu_s", C.c_float),
        ("microns_per_shell", C.c_float),
    ]

    def __repr__(self):
        return (f"Params(shells={self.shells}, mu_a={self.mu_a}, mu_s={self.mu_s}, "
                f"microns_per_shell={self.microns_per_shell})")


class Scales(C.Structure):
    _fields_ = [
        ("heat_shift", C.c_uint32),
        ("heat2_rshift", C.c_uint32),
        ("absorb_q32", C.c_uint32),
        ("roulette_thr", C.c_uint32),
    ]


class RunInfo(C.Structure):
    _fields_ = [
        ("photons", C.c_uint64),
        ("events", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("call_ms", C.c_double),
        ("n_gpus", C.c_uint32),
        ("gpu_launches", C.c_uint32),
        ("blocks_per_gpu", C.c_uint32),
        ("threads_per_block", C.c_uint32),
        ("philox_rounds", C.c_uint32),
        ("flush_iters", C.c_uint32),
        ("smem_bytes", C.c_uint32),
        ("retries", C.c_uint32),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# The named configurations of BASELINE.json (configs[0..4]).
CONFIGS = {
    "default": dict(shells=101, mu_a=2.0, mu_s=20.0, microns_per_shell=50.0),
    "highalbedo": dict(shells=101, mu_a=0.1, mu_s=100.0, microns_per_shell=50.0),
    "finegrid": dict(shells=16384, mu_a=2.0, mu_s=20.0, microns_per_shell=5.0),
}

# every symbol include/tiny_mc_b200.h declares
EXPORTS = (
    "tmc_init", "tmc_prepare", "tmc_finalize", "tmc_device_count", "tmc_last_error", "tmc_version",
    "tmc_abi_version", "tmc_set_option", "tmc_photons", "tmc_photons_fx", "tmc_photons_device",
    "tmc_fx_scales", "tmc_fx_accumulate", "tmc_generation_plan", "tmc_last_run_info",
    "tmc_photons_fx_batches", "tmc_device_tallies_check",
)

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("TMC_LIB", _ROOT / "lib" / "libtinymc_b200.so"))


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise TinyMcError(-1, f"{path} not built: run `make lib` or __graft_entry__.build() (no CPU fallback)")
    lib = C.CDLL(str(path))
    u64, p = C.c_uint64, C.c_void_p
    lib.tmc_init.argtypes = [C.c_int]
    lib.tmc_prepare.argtypes = [C.POINTER(Params)]
    lib.tmc_set_option.argtypes = [C.c_char_p, C.c_longlong]
    lib.tmc_last_error.restype = C.c_char_p
    lib.tmc_version.restype = C.c_char_p
    lib.tmc_photons.argtypes = [C.POINTER(Params), u64, u64, u64, p, p]
    lib.tmc_photons_fx.argtypes = [C.POINTER(Params), u64, u64, u64, p, p]
    lib.tmc_photons_device.argtypes = [C.POINTER(Params), u64, u64, u64, C.c_int, p, p]
    lib.tmc_photons_fx_batches.argtypes = [C.POINTER(Params), u64, u64, u64, C.c_uint32, p, p]
    lib.tmc_device_tallies_check.argtypes = [C.POINTER(Params), C.c_int, p, p]
    lib.tmc_fx_scales.argtypes = [C.POINTER(Params), C.POINTER(Scales)]
    lib.tmc_fx_accumulate.argtypes = [C.POINTER(Params), p, p, p, p]
    lib.tmc_last_run_info.argtypes = [C.POINTER(RunInfo)]
    lib.tmc_generation_plan.argtypes = [C.POINTER(Params), C.c_uint32, p, p, p]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise TinyMcError(rc, load().tmc_last_error().decode())


def make_params(cfg) -> Params:
    if isinstance(cfg, Params):
        return cfg
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    return Params(int(cfg["shells"]), float(cfg["mu_a"]), float(cfg["mu_s"]), float(cfg["microns_per_shell"]))


def init(n_gpus: int = 0):
    _check(load().tmc_init(int(n_gpus)))
    return load().tmc_device_count()


def prepare(cfg):
    """Pay the one-off costs of a configuration outside any timed region."""
    p = make_params(cfg)
    _check(load().tmc_prepare(C.byref(p)))


def finalize():
    _check(load().tmc_finalize())


def set_option(name: str, value: int):
    _check(load().tmc_set_option(name.encode(), int(value)))


def photons(cfg, seed: int, first_photon: int, n_photons: int, heats: np.ndarray, heats_squared: np.ndarray):
    """Batched photon(): ADD photons [first, first+n) into the caller's float32[SHELLS] tallies."""
    p = make_params(cfg)
    for a in (heats, heats_squared):
        if a.dtype != np.float32 or a.shape != (p.shells,) or not a.flags.c_contiguous:
            raise ValueError("tallies must be contiguous float32[SHELLS]")
    _check(load().tmc_photons(C.byref(p), seed, first_photon, n_photons, heats.ctypes.data, heats_squared.ctypes.data))


def photons_fx(cfg, seed: int, first_photon: int, n_photons: int, heat_fx: np.ndarray = None, heat2_fx: np.ndarray = None):
    """Exact fixed-point tallies (uint64[SHELLS] each), added into the given arrays or fresh zeros."""
    p = make_params(cfg)
    if heat_fx is None:
        heat_fx = np.zeros(p.shells, np.uint64)
    if heat2_fx is None:
        heat2_fx = np.zeros(p.shells, np.uint64)
    for a in (heat_fx, heat2_fx):
        if a.dtype != np.uint64 or a.shape != (p.shells,) or not a.flags.c_contiguous:
            raise ValueError("fixed-point tallies must be contiguous uint64[SHELLS]")
    _check(load().tmc_photons_fx(C.byref(p), seed, first_photon, n_photons, heat_fx.ctypes.data, heat2_fx.ctypes.data))
    return heat_fx, heat2_fx


def photons_device(cfg, seed: int, first_photon: int, n_photons: int, device: int, d_tallies_ptr: int, stream_ptr: int = 0):
    """Asynchronous device-resident form: add into a device uint64[2*SHELLS+4] buffer."""
    p = make_params(cfg)
    _check(load().tmc_photons_device(C.byref(p), seed, first_photon, n_photons, int(device), C.c_void_p(d_tallies_ptr), C.c_void_p(stream_ptr)))


def photons_fx_batches(cfg, seed: int, first_photon: int, n_photons: int, n_batches: int):
    """n_batches consecutive sub-ranges in one pass: (heat_fx, heat2_fx) as uint64[n_batches, SHELLS]."""
    p = make_params(cfg)
    heat_fx = np.zeros((n_batches, p.shells), np.uint64)
    heat2_fx = np.zeros((n_batches, p.shells), np.uint64)
    _check(load().tmc_photons_fx_batches(C.byref(p), seed, first_photon, n_photons, n_batches, heat_fx.ctypes.data, heat2_fx.ctypes.data))
    return heat_fx, heat2_fx


def device_tallies_check(cfg, device: int, d_tallies_ptr: int, stream_ptr: int = 0):
    """Raise TinyMcError(5) when the range flag of a device tally buffer is set (mandatory after photons_device)."""
    p = make_params(cfg)
    _check(load().tmc_device_tallies_check(C.byref(p), int(device), C.c_void_p(d_tallies_ptr), C.c_void_p(stream_ptr)))


def fx_scales(cfg) -> Scales:
    p = make_params(cfg)
    s = Scales()
    _check(load().tmc_fx_scales(C.byref(p), C.byref(s)))
    return s


def fx_accumulate(cfg, heat_fx: np.ndarray, heat2_fx: np.ndarray, heats: np.ndarray, heats_squared: np.ndarray):
    p = make_params(cfg)
    _check(load().tmc_fx_accumulate(C.byref(p), heat_fx.ctypes.data, heat2_fx.ctypes.data, heats.ctypes.data, heats_squared.ctypes.data))


def generation_plan(cfg, max_gen: int = 8):
    """(first_event, n_events, w_start) per generation of the deterministic weight schedule."""
    prm = make_params(cfg)
    a, b, c = (np.zeros(max_gen, np.uint32) for _ in range(3))
    n = load().tmc_generation_plan(C.byref(prm), max_gen, a.ctypes.data, b.ctypes.data, c.ctypes.data)
    if n < 0:
        _check(-n)
    return a[:n], b[:n], c[:n]


def fx_to_float64(cfg, heat_fx: np.ndarray, heat2_fx: np.ndarray):
    """heat, heat2 in weight units as float64 (for statistics; not part of the C ABI)."""
    s = fx_scales(cfg)
    heat = heat_fx.astype(np.float64) * 2.0 ** -int(s.heat_shift)
    heat2 = heat2_fx.astype(np.float64) * 2.0 ** (int(s.heat2_rshift) - 2 * int(s.heat_shift))
    return heat, heat2


def last_run_info() -> RunInfo:
    info = RunInfo()
    _check(load().tmc_last_run_info(C.byref(info)))
    return info
