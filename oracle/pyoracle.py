"""oracle/pyoracle.py — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

ctypes access to liboracle.so (port of reference photon.c + stream replay + batch harness)
and, when present, to oracle/_ref (the unmodified reference object code).  Imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import multiprocessing as mp
import os
import subprocess
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"

CONFIGS = {
    "default": dict(shells=101, mu_a=2.0, mu_s=20.0, microns_per_shell=50.0),
    "highalbedo": dict(shells=101, mu_a=0.1, mu_s=100.0, microns_per_shell=50.0),
    "finegrid": dict(shells=16384, mu_a=2.0, mu_s=20.0, microns_per_shell=5.0),
}


class Optics(C.Structure):
    _fields_ = [("shells", C.c_uint32), ("mu_a", C.c_float), ("mu_s", C.c_float), ("microns_per_shell", C.c_float)]


class FxScales(C.Structure):
    _fields_ = [("weight_one", C.c_uint32), ("heat_shift", C.c_uint32), ("heat2_rshift", C.c_uint32),
                ("absorb_q32", C.c_uint32), ("roulette_thr", C.c_uint32)]


def optics(cfg) -> Optics:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    return Optics(int(cfg["shells"]), float(cfg["mu_a"]), float(cfg["mu_s"]), float(cfg["microns_per_shell"]))


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(HERE), "all"], check=True)
    if Path("/root/reference/photon.c").exists():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = HERE / "liboracle.so"
        if not path.exists():
            build()
        l = C.CDLL(str(path))
        l.orc_photon.restype = C.c_uint32
        l.orc_photon.argtypes = [C.POINTER(Optics), C.c_void_p, C.c_void_p]
        l.orc_run_batch.restype = C.c_uint64
        l.orc_run_batch.argtypes = [C.POINTER(Optics), C.c_void_p, C.c_int, C.c_uint, C.c_uint64, C.c_uint32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_philox4x32.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_fx_plan.argtypes = [C.POINTER(Optics), C.POINTER(FxScales)]
        l.orc_replay.restype = C.c_uint64
        l.orc_replay.argtypes = [C.POINTER(Optics), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        l.orc_replay_mode.restype = C.c_uint64
        l.orc_replay_mode.argtypes = [C.POINTER(Optics), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        l.orc_replay_per_photon.restype = C.c_uint64
        l.orc_replay_per_photon.argtypes = [C.POINTER(Optics), C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_step_of_word.restype = C.c_float
        l.orc_step_of_word.argtypes = [C.c_uint32]
        l.orc_xi_of_word.restype = C.c_double
        l.orc_xi_of_word.argtypes = [C.c_uint32]
        l.orc_step_moments.argtypes = [C.c_void_p]
        l.orc_costheta_of_word.restype = C.c_float
        l.orc_costheta_of_word.argtypes = [C.c_uint32]
        l.orc_generation_plan.restype = C.c_uint32
        l.orc_generation_plan.argtypes = [C.POINTER(Optics), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


def have_ref(name: str = "default") -> bool:
    return (REF_DIR / f"libphoton_{name}.so").exists()


_ref_cache = {}


def ref_photon_ptr(name: str) -> C.c_void_p:
    """Address of the UNMODIFIED reference photon() compiled for the named configuration."""
    if name not in _ref_cache:
        _ref_cache[name] = C.CDLL(str(REF_DIR / f"libphoton_{name}.so"))
    return C.cast(_ref_cache[name].photon, C.c_void_p)


def ref_pcg_lib(name: str) -> C.CDLL:
    """The UNMODIFIED reference photon.c compiled with -Drand=pcg31 (oracle/Makefile, oracle/pcg31.c)."""
    key = "pcg_" + name
    if key not in _ref_cache:
        l = C.CDLL(str(REF_DIR / f"libphoton_pcg_{name}.so"))
        l.pcg31_seed.argtypes = [C.c_uint64]
        _ref_cache[key] = l
    return _ref_cache[key]


RNG_KINDS = {"libc": 0, "xoshiro": 1, "pcg": 2}


def run_batch(cfg, seed: int, n_photons: int, chunk: int = 256, impl: str = "port", ref_name: str = None, rng: str = "libc"):
    """One srand(seed) batch.  impl: "reference" (oracle/_ref object code, libc rand() only),
    "reference_pcg" (the same unmodified source with rand() bound to PCG32 at compile time) or
    "port" (photon_port.c; rng "libc" = the reference's stream, "xoshiro" = a sound generator).  Returns dict(heat, heat2 [float64], heat_f, heat2_f [float32 when chunk==0], events)."""
    o = optics(cfg)
    heat = np.zeros(o.shells)
    heat2 = np.zeros(o.shells)
    heat_f = np.zeros(o.shells, np.float32)
    heat2_f = np.zeros(o.shells, np.float32)
    fn = None
    if impl == "reference":
        fn = ref_photon_ptr(ref_name or (cfg if isinstance(cfg, str) else "default"))
    elif impl == "reference_pcg":
        l = ref_pcg_lib(ref_name or (cfg if isinstance(cfg, str) else "default"))
        l.pcg31_seed(seed)
        fn = C.cast(l.photon, C.c_void_p)
    t0 = time.perf_counter()
    ev = lib().orc_run_batch(C.byref(o), fn, RNG_KINDS[rng], seed, n_photons, chunk, heat.ctypes.data, heat2.ctypes.data,
                             heat_f.ctypes.data, heat2_f.ctypes.data)
    dt = time.perf_counter() - t0
    return dict(heat=heat, heat2=heat2, heat_f=heat_f, heat2_f=heat2_f, events=int(ev), seconds=dt)


def _batch_worker(args):
    cfg, seed, n, chunk, impl, ref_name, rng = args
    r = run_batch(cfg, seed, n, chunk, impl, ref_name, rng)
    return r["heat"], r["heat2"], r["events"], r["seconds"]


def run_batches(cfg, seeds, n_per_batch: int, chunk: int = 256, impl: str = "port", ref_name: str = None, processes: int = None,
                rng: str = "libc"):
    """Independent batches (distinct srand seeds) fanned out over host cores as separate
    PROCESSES — libc rand() is process-global state, so threads would share one stream.
    Returns (heat[B, S], heat2[B, S], events[B], seconds[B], wall_seconds)."""
    processes = processes or min(len(seeds), os.cpu_count() or 1)
    jobs = [(cfg, int(s), int(n_per_batch), chunk, impl, ref_name, rng) for s in seeds]
    t0 = time.perf_counter()
    if processes <= 1:
        out = [_batch_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(processes) as pool:
            out = pool.map(_batch_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    heat = np.stack([o[0] for o in out])
    heat2 = np.stack([o[1] for o in out])
    return heat, heat2, np.array([o[2] for o in out]), np.array([o[3] for o in out]), wall


def philox4x32(rounds: int, ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32(rounds, c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def fx_plan(cfg) -> FxScales:
    s = FxScales()
    o = optics(cfg)
    lib().orc_fx_plan(C.byref(o), C.byref(s))
    return s


def replay(cfg, seed: int, first: int, n: int, rounds: int = 10, mode: int = 0):
    """CPU replay of the product's Philox stream: exact u64 fixed-point tallies + event count.
    mode 0 = the 3-D walk, 1 = the reduced radial cross-check walk."""
    o = optics(cfg)
    heat = np.zeros(o.shells, np.uint64)
    heat2 = np.zeros(o.shells, np.uint64)
    ev = lib().orc_replay_mode(C.byref(o), rounds, seed, first, n, mode, heat.ctypes.data, heat2.ctypes.data)
    return heat, heat2, int(ev)


def replay_per_photon(cfg, seed: int, first: int, n: int, rounds: int = 10):
    """Replay with the per-PHOTON second moment: returns (heat, per_photon_sq) in weight units (float64[shells]);
    Var(mean heat[s]) = (per_photon_sq[s] / n - (heat[s] / n)^2) / n."""
    o = optics(cfg)
    heat = np.zeros(o.shells, np.uint64)
    heat2 = np.zeros(o.shells, np.uint64)
    sq = np.zeros(o.shells, np.float64)
    lib().orc_replay_per_photon(C.byref(o), rounds, seed, first, n, heat.ctypes.data, heat2.ctypes.data, sq.ctypes.data)
    return fx_to_float64(cfg, heat, heat2)[0], sq


def generation_plan(cfg, max_gen: int = 8):
    """(first_event, n_events, w_start) per generation of the deterministic weight schedule."""
    o = optics(cfg)
    a, b, c = (np.zeros(max_gen, np.uint32) for _ in range(3))
    lib().orc_generation_plan(C.byref(o), max_gen, a.ctypes.data, b.ctypes.data, c.ctypes.data)
    return a, b, c


def fx_to_float64(cfg, heat_fx, heat2_fx):
    s = fx_plan(cfg)
    return (heat_fx.astype(np.float64) * 2.0 ** -int(s.heat_shift),
            heat2_fx.astype(np.float64) * 2.0 ** (int(s.heat2_rshift) - 2 * int(s.heat_shift)))
