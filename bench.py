#!/usr/bin/env python
"""bench.py — photons/s of the photon random-walk hot path on 1..8 B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config default]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A "step" is one pass of the hot path (reference tiny_mc.c:47-49, i.e. PHOTONS calls of photon())
over one batch of photons:
  N = 1 : BASELINE.json configs[1] — default optics, 2^26 photons per step on one GPU;
  N > 1 : 2^29 photons per GPU per step (weak scaling; N = 8 is exactly configs[2]: 2^32 photons
          sharded over 8 GPUs), one NCCL all-reduce of the 2*SHELLS+4 u64 tally words per step.
There is no input data: the "inputs" are the photon index range and the seed, so nothing has
to be resident in HBM and nothing is copied host->device; every step simulates a NEW photon range.

  value : device-resident — tmc_photons_device() into a device tally buffer on torch's current
          stream, CUDA events on that stream, max over ranks.
  e2e   : the reference-facing C-ABI call tmc_photons() with HOST float tallies (kernel + D2H of
          the tally words + float accumulation on the host), wall clock (N = 1: the library drives
          the GPU itself; N > 1: device call + all-reduce + D2H per step).
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
MUFU_PER_EVENT_CANONICAL = 3.0     # lg2, sqrt, sqrt (simplified Marsaglia)
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
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tiny_mc_b200 as tmc

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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = tmc.CONFIGS[args.config]
    shells = cfg["shells"]
    per_gpu = args.photons_per_gpu or ((1 << 26) if world == 1 else (1 << 29))
    if args.config == "highalbedo" and not args.photons_per_gpu:
        per_gpu >>= 4    # 2^22 photons = 3e10 events per step: ~28 cohorts per warp, so the one-cohort tail is ~1 %
    per_step = per_gpu * world
    seed = 0x5EED
    tmc.set_option("philox_rounds", args.philox_rounds)
    words = 2 * shells + 4
    tallies = torch.zeros(words, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from tiny_mc_b200.shards import shard_range

    def device_step(step: int):
        first, count = shard_range(step * per_step, per_step, rank, world)    # this rank's photon indices
        tmc.photons_device(args.config, seed, first, count, local_rank, tallies.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(tallies)     # the single collective: 2*SHELLS+4 int64 words, exact

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for w in range(args.warmup):
        tallies.zero_()
        device_step(w)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    tallies.zero_()
    total_tally_check = 0
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sync_all()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush_buf.fill_(k & 0xFF)       # L2 flush between timed iterations (outside the events)
        if world > 1:
            tallies.zero_()              # every rank contributes its own step tallies to the reduce
        starts[k].record(stream)
        device_step(args.warmup + k)
        stops[k].record(stream)
    sync_all()
    wall = time.perf_counter() - wall0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    # the same loop with Philox4x32-7 (the smallest Crush-resistant variant of Salmon et al.), reported beside the headline
    philox7_value = None
    if args.philox_rounds == 10 and world == 1:
        tmc.set_option("philox_rounds", 7)
        for w in range(args.warmup):
            device_step(w)
        s7 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        s7[0].record(stream)
        for k in range(args.steps):
            device_step(args.warmup + k)
        s7[1].record(stream)
        torch.cuda.synchronize()
        philox7_value = per_step * args.steps / (s7[0].elapsed_time(s7[1]) * 1e-3)
        tmc.set_option("philox_rounds", 10)
    clocks = sampler.stop() if rank == 0 else None
    host_tallies = tallies.cpu().numpy().astype(np.uint64)
    events_last = int(host_tallies[2 * shells])
    flag = int(host_tallies[2 * shells + 2])
    photons_counted = int(host_tallies[2 * shells + 1])
    if world == 1:
        events_per_photon = events_last / max(photons_counted, 1)
    else:
        events_per_photon = events_last / max(photons_counted, 1)
    value = per_step * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the reference-facing C-ABI call with HOST tallies ("e2e") ----
    heat = np.zeros(shells, np.float32)
    heat2 = np.zeros(shells, np.float32)
    d2h = words * 8
    if world == 1:
        tmc.init(1)
        for w in range(args.warmup):
            tmc.photons(args.config, seed, w * per_step, per_gpu, heat, heat2)
        heat[:] = 0
        heat2[:] = 0
        torch.cuda.synchronize()
        e0 = time.perf_counter()
        for k in range(args.steps):
            tmc.photons(args.config, seed, (args.warmup + k) * per_step, per_gpu, heat, heat2)
        e2e_s = time.perf_counter() - e0
        info = tmc.last_run_info().as_dict()
        absorbed = float(heat.sum()) / (per_gpu * args.steps)
    else:
        pinned = torch.empty(words, dtype=torch.int64).pin_memory()
        sync_all()
        e0 = time.perf_counter()
        for k in range(args.steps):
            tallies.zero_()
            device_step(args.warmup + k)
            pinned.copy_(tallies, non_blocking=False)       # D2H of the reduced tally words
            fx = pinned.numpy().astype(np.uint64)
            tmc.fx_accumulate(args.config, fx[:shells].copy(), fx[shells:2 * shells].copy(), heat, heat2)
        sync_all()
        e2e_s = time.perf_counter() - e0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        info = tmc.last_run_info().as_dict()
        absorbed = float(heat.sum()) / (per_step * args.steps)
    e2e_value = per_step * args.steps / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: issue slots (FP32/INT) and MUFU, SURVEY §8(d) ----
    f_sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    events_per_s_per_gpu = value / world * events_per_photon
    issue_peak = sms * LANES_PER_CLK_PER_SM * f_sm_hz
    mufu_peak = sms * MUFU_PER_CLK_PER_SM * f_sm_hz
    achieved = events_per_s_per_gpu * CANONICAL_SLOTS_PER_EVENT
    roofline = {
        "bound": "issue",
        "achieved": achieved / 1e12, "peak": issue_peak / 1e12, "unit": "T lane-instr/s",
        "frac": achieved / issue_peak,
        "traffic": 79104,   # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu capture profiles/r01_v5_default_ncu.md
        "definition": "events/s/GPU x 88 canonical issue slots per event (SURVEY 8d: walk 35 + Philox4x32-10 53) "
                      "/ (SMs x 128 lanes/clk x SM clock sampled during the run)",
        "frac_walk_only_35_slots": events_per_s_per_gpu * WALK_SLOTS_PER_EVENT / issue_peak,
        "frac_mufu_3_per_event": events_per_s_per_gpu * MUFU_PER_EVENT_CANONICAL / mufu_peak,
        "events_per_s_per_gpu": events_per_s_per_gpu,
        "events_per_photon": events_per_photon,
        "peak_source": f"{sms} SMs x 128 lanes/clk x {f_sm_hz / 1e6:.0f} MHz (nvidia-smi median under load); "
                       "MEASURED_PEAKS.json holds HBM and bf16 peaks only - the issue rate was measured on the box: "
                       "127.4 lanes/clk/SM sustained (profiles/r01_microbench_pipes.md)",
        "hardware_truth": "ncu of this kernel (profiles/): warp instructions per event, issue-slot, FMA-heavy, ALU, XU and "
                          "shared-pipe utilisation; frac > 1 means fewer instructions than the canonical 88-slot budget",
        "philox10_ceiling_events_per_s": sms * 4 * f_sm_hz / 80.0 * 32 * 3,
        "frac_of_philox10_ceiling": events_per_s_per_gpu / (sms * 4 * f_sm_hz / 80.0 * 32 * 3),
        "bytes_note": "HBM traffic is ~80 KB per launch (ncu dram__bytes): the kernel reads no input",
    }

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+u32 (fp32 walk, 32-bit fixed-point weights, u64 tallies)", "data": "synthetic",
        "config": {
            "workload": (f"{args.config} optics (SHELLS={shells}, MU_A={cfg['mu_a']}, MU_S={cfg['mu_s']}, "
                         f"{cfg['microns_per_shell']} um shells), {per_gpu} photons per GPU per step, "
                         f"{per_step} photons per step" + (" = BASELINE configs[1]" if world == 1 and per_gpu == 1 << 26 else "")
                         + (" = BASELINE configs[2]" if per_step == 1 << 32 else "")),
            "parallelism": f"photon-index shards x{world}" + (", one NCCL all-reduce of the tally words per step" if world > 1 else ""),
            "philox_rounds": args.philox_rounds,
            "l2": "flushed between timed steps (256 MB fill); the kernel has no input to cache",
            "blocks": info["blocks_per_gpu"], "threads_per_block": info["threads_per_block"],
            "flush_iters": info["flush_iters"],
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                "note": "C-ABI tmc_photons() with host float tallies; no inputs exist to copy host->device "
                        "(launch arguments only); D2H = 2*SHELLS+4 u64 tally words"},
        "gpu_launches": args.steps * world,
        "roofline": roofline,
        "philox7": {"value": philox7_value, "unit": METRIC, "note": "same workload with philox_rounds=7; the headline uses Philox4x32-10"},
        "checks": {"absorbed_weight_per_photon": absorbed, "tally_range_flag": flag, "wall_ms_per_step": 1e3 * wall / args.steps},
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
