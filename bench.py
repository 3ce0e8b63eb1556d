#!/usr/bin/env python
"""bench.py — photons/s of the photon random-walk hot path on 1..8 B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config default]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A "step" is one pass of the hot path (reference tiny_mc.c:47-49, i.e. PHOTONS calls of photon())
over one batch of photons:
  N = 1 : BASELINE.json configs[1] — default optics, 2^26 photons per step on one GPU;
  N > 1 : 2^29 photons per GPU per step (weak scaling; N = 8 is exactly configs[2]: 2^32 photons
          sharded over 8 GPUs), one NCCL all-reduce of the 2*SHELLS+4 u64 tally words per step.
          `also` : the other named size of the same N — configs[2] (2^32 photons in total) at N = 2 and 4,
          the 2^29-photon step at N = 1 (the anchor that makes value_N / (N x anchor) compare equal work).
There is no input data: the "inputs" are the photon index range and the seed, so nothing has
to be resident in HBM and nothing is copied host->device; every step simulates a NEW photon range.

  value : device-resident — tmc_photons_device() into a device tally buffer on torch's current
          stream, CUDA events on that stream, max over ranks.
  e2e   : the reference-facing C-ABI call tmc_photons() with HOST float tallies, wall clock: ONE
          process drives all N GPUs (tmc_init(N): library streams, in-library ncclReduce, D2H of
          the tally words, float accumulation) — the path the C host program takes.  Under
          torchrun rank 0 runs it while the other ranks wait on the host (gloo barrier).
  checks.tally_hash : the tally words of one fixed photon range, sharded over the N GPUs —
          the same hash at every N (and through both host paths) is the bit-reproducibility proof.
  --impl reference : the UNMODIFIED reference photon() (oracle/_ref, compiled from
          /root/reference/photon.c) on all host cores as independent processes (libc rand() is
          process-global; the repo's variant has no OpenMP), bounded sample per step.

Prints ONE JSON line (rank 0).  The oracle is executed only in the cpu_baseline / reference legs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

_JSON_FD = None
ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "photons/s"
CANONICAL_SLOTS_PER_EVENT = 88.0   # walk 35 + canonical Philox4x32-10 53
WALK_SLOTS_PER_EVENT = 35.0        # RNG-free lower bound
LANES_PER_CLK_PER_SM = 128.0
MUFU_PER_CLK_PER_SM = 16.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="default", choices=["default", "highalbedo", "finegrid"])
    ap.add_argument("--photons-per-gpu", type=int, default=0, help="photons per GPU per step (0 = named config)")
    ap.add_argument("--philox-rounds", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-photons-per-core", type=int, default=0)
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=self.file, stderr=subprocess.DEVNULL)
        except FileNotFoundError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [r.split(", ") for r in Path(self.file.name).read_text().strip().splitlines() if r.strip()]
        os.unlink(self.file.name)
        sm, mx, power, reasons = [], [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU reference legs
def cpu_reference_run(config: str, photons_per_core: int, repeats: int = 1):
    """The reference's own CPU implementation on all host cores (oracle/_ref when it was built
    from /root/reference, else the bit-identical port).  Returns (photons/s, info dict)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import pyoracle as orc

    cores = os.cpu_count() or 1
    kind = "reference" if orc.have_ref(config) else "port"
    impl = "reference" if kind == "reference" else "port"
    best = 0.0
    for rep in range(repeats):
        seeds = [90001 + 131 * rep + c for c in range(cores)]
        _, _, _, secs, wall = orc.run_batches(config, seeds, photons_per_core, chunk=256, impl=impl, processes=cores)
        best = max(best, cores * photons_per_core / wall)
    info = {"value": best, "unit": METRIC, "cores": cores, "kind": kind,
            "sample": f"{cores} processes x {photons_per_core} photons ({config} optics), distinct srand seeds, "
                      f"gcc -O3 -march=x86-64-v3, libc rand(); no OpenMP in the reference",
            "single_core_photons_per_s": photons_per_core / float(max(secs))}
    return best, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_core = args.cpu_photons_per_core or ((1 << 16) if args.config != "highalbedo" else (1 << 10))
    for _ in range(args.warmup):
        cpu_reference_run(args.config, max(per_core // 8, 64))
    t0 = time.perf_counter()
    values = []
    for _ in range(args.steps):
        v, info = cpu_reference_run(args.config, per_core)
        values.append(v)
    wall = time.perf_counter() - t0
    cores = info["cores"]
    value = cores * per_core * args.steps / wall
    info["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} optics; reference photon() on {cores} host cores, "
                               f"{cores * per_core} photons per step (bounded sample of the GPU workload)"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ------------------------------------------------------------------------------ our arm
NAMED = {   # photons per step on one GPU (N = 1), photons per GPU per step on N > 1 GPUs (weak scaling)
    "default": ((1 << 26, "BASELINE configs[1]"), 1 << 29),          # 8 x 2^29 = 2^32 = configs[2]
    # 2^24 photons = 1.2e11 events per step on one GPU: ~110 cohorts per warp, so the one-cohort tail is ~1 % (2^22: 4 %)
    "highalbedo": ((1 << 24, "BASELINE configs[3] optics, 2^24-photon steps"), 1 << 27),     # 8 x 2^27 = 2^30 = configs[3]
    "finegrid": ((1 << 26, "BASELINE configs[4] optics, 2^26-photon steps"), 1 << 27),       # 8 x 2^27 = 2^30 = configs[4]
}
CONFIGS2_PHOTONS = 1 << 32                      # BASELINE configs[2]: 2^32 photons sharded over 2 / 4 / 8 GPUs
HASH_SEED, HASH_PHOTONS = 0x5EED, 1 << 26      # the fixed range behind checks.tally_hash at every N


def fnv64(words) -> str:
    """FNV-1a (64 bit) over the little-endian bytes of the u64 tally words."""
    h = 0xCBF29CE484222325
    for b in words.astype("<u8").tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def ncu_facts(config: str):
    """Per-event instruction counts of the shipped kernel from the committed ncu capture (tools/ncu_facts.py)."""
    path = ROOT / "profiles" / "r02_ncu_facts.json"
    try:
        return json.loads(path.read_text()).get(config)
    except (OSError, ValueError):
        return None


def mix_ceiling(instr_per_event: float, issue_peak: float, events_per_s: float):
    """The issue-slot utilisation a B200 SM reaches on the walk's own instruction mix with perfect instruction-level
    parallelism (tmc_microbench k_walk_mix under ncu, committed in profiles/r02_ncu_facts.json): no pipe is saturated
    there either; the half-rate ALU-pipe instructions and the two-slot IMAD.WIDE cap the mix at ~69 % of 128 lanes/clk."""
    path = ROOT / "profiles" / "r02_ncu_facts.json"
    try:
        m = json.loads(path.read_text())["_walk_mix_ceiling"]
    except (OSError, ValueError, KeyError):
        return None
    frac = m["issue_active_pct"] / 100.0
    return {"issue_frac_of_the_mix": frac, "frac_of_mix_ceiling": events_per_s * instr_per_event / issue_peak / frac,
            "source": "profiles/r02_microbench_walk_mix.md"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tiny_mc_b200 as tmc
    from tiny_mc_b200.shards import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tiny_mc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # host-side waits that keep every GPU idle

    cfg = tmc.CONFIGS[args.config]
    shells = cfg["shells"]
    if world == 1:
        per_step, named = NAMED[args.config][0]
    else:
        per_step = NAMED[args.config][1] * world
        if args.config != "default":
            named = f"weak scaling at {NAMED[args.config][1]} photons per GPU"
        elif per_step == CONFIGS2_PHOTONS:
            named = "BASELINE configs[2]"
        else:
            named = "weak-scaling step of BASELINE configs[2]: the per-GPU share of its 8-GPU run"
    if args.photons_per_gpu:
        per_step, named = args.photons_per_gpu * world, "custom size"
    per_gpu = per_step // world
    seed = 0x5EED
    tmc.set_option("philox_rounds", args.philox_rounds)
    words = 2 * shells + 4
    tallies = torch.zeros(words, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    launches = 0

    def device_walk(first_photon: int, n_photons: int):
        nonlocal launches
        first, count = shard_range(first_photon, n_photons, rank, world)    # this rank's photon indices
        tmc.photons_device(args.config, seed, first, count, local_rank, tallies.data_ptr(), stream.cuda_stream)
        launches += tmc.last_run_info().gpu_launches
        if world > 1:
            dist.all_reduce(tallies)     # the single collective: 2*SHELLS+4 int64 words, exact

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_steps(photons_per_step: int, warmup: int, steps: int, first_photon: int):
        """`warmup` untimed and `steps` timed passes over consecutive photon ranges of `photons_per_step`;
        returns (device ms of the timed steps, max over ranks; kernels launched in them, all ranks; wall s)."""
        nonlocal launches
        for w in range(warmup):
            tallies.zero_()
            device_walk(first_photon + w * photons_per_step, photons_per_step)
        sync_all()
        tallies.zero_()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        sync_all()
        launches = 0
        wall0 = time.perf_counter()
        for k in range(steps):
            flush_buf.fill_(k & 0xFF)       # L2 flush between timed iterations (outside the events)
            if world > 1:
                tallies.zero_()              # every rank contributes its own step tallies to the reduce
            starts[k].record(stream)
            device_walk(first_photon + (warmup + k) * photons_per_step, photons_per_step)
            stops[k].record(stream)
        sync_all()
        wall_s = time.perf_counter() - wall0
        ms = sum(s_.elapsed_time(e_) for s_, e_ in zip(starts, stops))
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
        n_launches = launches
        if world > 1:
            tl = t.clone()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(tl, op=dist.ReduceOp.SUM)
            n_launches = int(tl[1].item())
        return float(t[0].item()), n_launches, wall_s

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, timed_launches, wall = timed_steps(per_step, args.warmup, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    tmc.device_tallies_check(args.config, local_rank, tallies.data_ptr(), stream.cuda_stream)   # mandatory after photons_device
    host_tallies = tallies.cpu().numpy().astype(np.uint64)
    events_last = int(host_tallies[2 * shells])
    flag = int(host_tallies[2 * shells + 2])
    photons_counted = int(host_tallies[2 * shells + 1])
    events_per_photon = events_last / max(photons_counted, 1)
    value = per_step * args.steps / (dev_ms * 1e-3)

    # ---- the other named size of this N (default optics): N = 2, 4 also run BASELINE configs[2] (2^32 photons in
    #      total; at N = 8 the weak-scaling step above IS configs[2]); N = 1 also runs the 2^29-photon step of the
    #      N > 1 lines, so that value_N / (N x anchor) compares equal per-GPU work ----
    also = None
    if args.config == "default" and not args.photons_per_gpu:
        extra = NAMED["default"][1] if world == 1 else (CONFIGS2_PHOTONS if per_step != CONFIGS2_PHOTONS else 0)
        if extra:
            ms_x, _, _ = timed_steps(extra, 1, 3, 1 << 40)
            also = {"photons_per_step": extra, "value": extra * 3 / (ms_x * 1e-3), "unit": METRIC, "ms_per_step": ms_x / 3, "steps": 3,
                    "what": ("the 2^29-photon step the N > 1 lines run per GPU (weak-scaling anchor)" if world == 1
                             else f"BASELINE configs[2]: 2^32 photons per step sharded over {world} GPUs")}

    # ---- the same fixed photon range at every N: its tally words must hash the same (north star:
    #      "bit-reproducible across 1/2/4/8 GPUs"; reference loop tiny_mc.c:47-49 sharded) ----
    saved_seed, seed = seed, HASH_SEED
    tallies.zero_()
    device_walk(0, HASH_PHOTONS)
    torch.cuda.synchronize()
    seed = saved_seed
    hash_words = tallies.cpu().numpy().astype(np.uint64)[: 2 * shells]
    tally_hash = fnv64(hash_words)

    # ---- end to end through the reference-facing C-ABI call with HOST tallies ("e2e"): ONE process drives
    #      all N GPUs through tmc_init(N) + tmc_photons() (library streams, in-library NCCL reduce, D2H,
    #      float accumulation); under torchrun rank 0 does it while the other ranks wait on the host ----
    heat = np.zeros(shells, np.float32)
    heat2 = np.zeros(shells, np.float32)
    d2h = words * 8
    e2e_value = lib_hash = info = absorbed = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
    if rank == 0:
        tmc.init(world)
        tmc.prepare(args.config)
        for w in range(args.warmup):
            tmc.photons(args.config, seed, w * per_step, per_step, heat, heat2)
        heat[:] = 0
        heat2[:] = 0
        e0 = time.perf_counter()
        for k in range(args.steps):
            tmc.photons(args.config, seed, (args.warmup + k) * per_step, per_step, heat, heat2)
        e2e_s = time.perf_counter() - e0
        info = tmc.last_run_info().as_dict()
        absorbed = float(heat.astype(np.float64).sum()) / (per_step * args.steps)
        e2e_value = per_step * args.steps / e2e_s
        hfx, h2fx = tmc.photons_fx(args.config, HASH_SEED, 0, HASH_PHOTONS)
        lib_hash = fnv64(np.concatenate([hfx, h2fx]))
        tmc.finalize()
    if world > 1:
        dist.barrier(group=cpu_group)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: issue slots of the SM sub-partitions (FP32/INT/MUFU all issue through them), SURVEY §8(d) ----
    f_sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    events_per_s_per_gpu = value / world * events_per_photon
    issue_peak = sms * LANES_PER_CLK_PER_SM * f_sm_hz
    mufu_peak = sms * MUFU_PER_CLK_PER_SM * f_sm_hz
    facts = ncu_facts(args.config)
    roofline = {
        "bound": "issue",
        "unit": "T lane-instr/s",
        "peak": issue_peak / 1e12,
        "peak_source": f"{sms} SMs x 128 lanes/clk x {f_sm_hz / 1e6:.0f} MHz (nvidia-smi median under load); "
                       "MEASURED_PEAKS.json holds HBM and bf16 peaks only - the issue rate was measured on the box: "
                       "127.4 lanes/clk/SM sustained (profiles/r01_microbench_pipes.md)",
        "events_per_s_per_gpu": events_per_s_per_gpu,
        "events_per_photon": events_per_photon,
        "frac_canonical_88": events_per_s_per_gpu * CANONICAL_SLOTS_PER_EVENT / issue_peak,
        "frac_walk_only_35_slots": events_per_s_per_gpu * WALK_SLOTS_PER_EVENT / issue_peak,
        "frac_mufu_2_per_event": events_per_s_per_gpu * 2.0 / mufu_peak,
    }
    if facts:
        ncu_events_per_s = facts["events"] / (facts["duration_ms"] * 1e-3)
        roofline.update({
            "achieved": events_per_s_per_gpu * facts["warp_instr_per_event"] / 1e12,
            "frac": events_per_s_per_gpu * facts["warp_instr_per_event"] / issue_peak,
            "definition": "live events/s/GPU x warp instructions per event EXECUTED by this kernel (ncu smsp__inst_executed of the "
                          "committed capture, lane-normalised) / (SMs x 128 lanes/clk x SM clock sampled during the run) "
                          "= the hardware issue-slot utilisation (ncu sm__issue_active of the capture: "
                          f"{facts['issue_active_pct']:.1f} %)",
            "warp_instr_per_event": facts["warp_instr_per_event"],
            "frac_dispatch": (events_per_s_per_gpu * facts["dispatch_slots_per_event"] / issue_peak) if facts.get("dispatch_slots_per_event") else None,
            "frac_fma_heavy": facts["fma_heavy_pct"] / 100.0 * events_per_s_per_gpu / ncu_events_per_s,
            "frac_shared_pipe": facts["shared_pipe_pct"] / 100.0 * events_per_s_per_gpu / ncu_events_per_s,
            "mix_ceiling": mix_ceiling(facts["warp_instr_per_event"], issue_peak, events_per_s_per_gpu),
            "traffic": facts["dram_bytes"],
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full): the path reads no input, HBM is idle",
            "ncu_capture": {"file": "profiles/r02_ncu_facts.json", "report": facts["report"], "block_threads": facts["block_threads"],
                            "matches_this_run": facts["block_threads"] == info["threads_per_block"]},
        })
    else:
        roofline.update({"achieved": None, "frac": None, "traffic": None,
                         "definition": "no committed ncu capture for this configuration (profiles/r02_ncu_facts.json)"})

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32+u32 (fp32 walk, 32-bit fixed-point weights, u64 tallies)", "data": "synthetic",
        "config": {
            "workload": (f"{args.config} optics (SHELLS={shells}, MU_A={cfg['mu_a']}, MU_S={cfg['mu_s']}, "
                         f"{cfg['microns_per_shell']} um shells), {per_step} photons per step ({named}), "
                         f"{per_gpu} photons per GPU per step"),
            "parallelism": f"photon-index shards x{world}" + (", one NCCL all-reduce of the tally words per step" if world > 1 else ""),
            "scaling_note": "N = 1 is configs[1] (2^26 photons per step); N = 2, 4, 8 run 2^29 photons per GPU per step (weak "
                            "scaling; N = 8 is exactly configs[2]); `also` carries the other named size of this N: configs[2] "
                            "(2^32 photons in total) at N = 2 and 4, the 2^29-photon anchor at N = 1",
            "philox_rounds": args.philox_rounds,
            "l2": "flushed between timed steps (256 MB fill); the kernel has no input to cache",
            "blocks": info["blocks_per_gpu"], "threads_per_block": info["threads_per_block"],
            "flush_iters": info["flush_iters"],
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                "note": f"one process, tmc_init({world}) + C-ABI tmc_photons() with HOST float tallies per step: launches on "
                        f"{world} GPU(s)" + (", in-library ncclReduce," if world > 1 else ",") + " D2H of 2*SHELLS+4 u64 tally words, "
                        "float accumulation; wall clock.  No inputs exist to copy host->device (launch arguments only)",
                "library_kernel_ms_last_step": info["kernel_ms"], "library_call_ms_last_step": info["call_ms"]},
        "gpu_launches": timed_launches,
        "also": also,
        "roofline": roofline,
        "checks": {"absorbed_weight_per_photon": absorbed, "tally_range_flag": flag, "wall_ms_per_step": 1e3 * wall / args.steps,
                   "tally_hash": tally_hash, "tally_hash_library_path": lib_hash,
                   "tally_hash_of": f"FNV-1a-64 of heat_fx|heat2_fx (2*SHELLS u64 words), seed {HASH_SEED:#x}, photons [0, 2^26), "
                                    f"sharded over {world} GPU(s): identical at every N"},
    }
    if world == 1 and not args.no_cpu_baseline:
        per_core = args.cpu_photons_per_core or ((1 << 18) if args.config != "highalbedo" else (1 << 11))
        _, cpu = cpu_reference_run(args.config, per_core)
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: NCCL and torchrun print banners ("NCCL version ...") on fd 1,
    # so everything else is sent to stderr and the line is written to the saved descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
