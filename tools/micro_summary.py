#!/usr/bin/env python
"""tools/micro_summary.py — turn the ncu metrics of tmc_microbench (gpurun_out/micro_ncu.csv) into a table.

    python tools/micro_summary.py gpurun_out/micro_ncu.csv > profiles/r01_microbench_pipes.md
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
data = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    rec = dict(zip(hdr, r))
    data.setdefault((rec["Kernel Name"].split("(")[0], rec["ID"]), {})[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
last = collections.OrderedDict()
for (name, _), m in data.items():
    last[name] = m
print("# Pipe micro-benchmarks on the B200 (tiny_mc_b200/csrc/microbench.cu under `ncu --clock-control none`)\n")
print("148 SMs x 4 sub-partitions; every kernel runs 32 warps per SM for 16384 iterations of 16-24 independent")
print("instructions; rates are `smsp__inst_executed.sum / sm__cycles_elapsed.avg / 148` (loop overhead included),")
print("pipe columns are ncu's own utilisation counters.  `k_imad_wide` was folded by ptxas into 32-bit IMADs")
print("(its high half is dead) - the IMAD.WIDE.U32 rate is the one in `k_imad_wide_live` (Philox-shaped: both halves of")
print("every product are consumed) and `k_philox10` (20 IMAD.WIDE + 20 LOP3 per call).\n")
print("| kernel | warp-inst / clk / SM | FMA-heavy busy % | FMA (heavy+lite) busy % | ALU % | XU % | issue % | shared wavefronts / clk / SM |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for name, m in last.items():
    cyc = m["sm__cycles_elapsed.avg"]
    print(f"| `{name}` | {m['smsp__inst_executed.sum'] / cyc / 148:.3f} | {m['sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | "
          f"{m['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | {m['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']:.1f} | "
          f"{m['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']:.1f} | {m['sm__issue_active.avg.pct_of_peak_sustained_elapsed']:.1f} | "
          f"{m['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'] / cyc / 148:.3f} |")
print("""
Reading (these are the denominators DESIGN.md §7 uses):
* issue: 3.98 warp-instructions / clk / SM sustained with scalar FFMA/FMUL/FADD = 127 lane-instructions / clk / SM
  (nominal 128): the FP32 / issue roofline is 148 x 128 x f_SM.
* scalar FP32 splits over FMA-heavy and FMA-lite; FFMA2 (packed f32x2), IMAD and LOP3/SHF/PRMT/I2FP run at
  2 warp-instructions / clk / SM (one 16-lane pipe each: FMA-heavy resp. ALU).  FFMA2 runs ONLY on FMA-heavy:
  same FP32 throughput as scalar code at half the issue slots.
* IMAD.WIDE.U32 / IMAD.HI: ~4 heavy-pipe cycles per warp instruction (k_imad_wide_live, k_philox10: 0.26 IMAD.WIDE /
  clk / SMSP at 94-97 % heavy-pipe busy) - a Philox4x32-10 block costs ~80 heavy-pipe cycles per warp.
* an IMAD.WIDE costs what TWO 32-bit IMADs cost: a pipe-balanced group {IMAD.WIDE, 2 LOP3, 2 FMUL} takes 7.25 cycles
  per SMSP, the same group with {IMAD, IMAD} instead of the IMAD.WIDE 7.0 (k_dispatch_*); likewise FFMA2 + LOP3 pairs
  reach 0.73 instructions / clk / SMSP with neither pipe saturated (k_mix_ffma2_lop3).  DESIGN.md §7 therefore counts
  instructions with 64-bit register operands as two dispatch slots.
* MUFU (lg2, sqrt, rsq, ex2, sin) and F2I: 0.5 warp-instructions / clk / SM = 16 lanes / clk / SM.
* shared-memory atomics: ~1 wavefront / clk / SM; random 101-bin RED costs ~3.2 wavefronts per instruction,
  the lane-private layout exactly 1.
""")
