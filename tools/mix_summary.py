#!/usr/bin/env python
"""tools/mix_summary.py gpurun_out/mix_ncu.csv — the walk-mix ceiling micro-benchmarks (tmc_microbench: k_walk_mix<DROP>)
under ncu: instructions and cycles per 16-event trip, issue-slot and pipe utilisation."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr, data = rows[0], {}
for r in rows[1:]:
    d = dict(zip(hdr, r))
    data.setdefault(d["Kernel Name"].split("(")[0], {})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
names = {"k_walk_mix<0>": "the walk's mix", "k_walk_mix<1>": "without the 3 PRMT + SHF", "k_walk_mix<2>": "without the Philox quarter block",
         "k_walk_mix<3>": "without LDS / RED", "k_walk_mix<4>": "without the 2 MUFU", "k_walk_mix<5>": "without the FP32 arithmetic", "k_walk_mix<6>": "kernel v9's mix: packed f32x2 forms (FFMA2 for (y, z), FADD2 / FFMA2.RZ per two chains)",
         "k_walk_mix_32warps": "the walk's mix at 32 warps per SM (4 chains per thread, +0.5 LOP3 per event)"}
print("| kernel | what | instr / event | cycles / event / SMSP | issue slots used % | FMA-heavy % | ALU % | XU % | shared wavefronts / clk / SM |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
for k, m in data.items():
    cyc, inst = m["sm__cycles_elapsed.avg"], m["smsp__inst_executed.sum"]
    trips = 4096.0                                   # kIters / 4, 16 events each
    warps = 32.0 if "32warps" in k else 16.0         # per SM
    k = k.replace("void ", "")
    print(f"| `{k}` | {names.get(k, '')} | {inst / (148 * warps * trips * 16):.2f} | {cyc / (trips * (warps / 4) * 16):.2f} | "
          f"{m['sm__issue_active.avg.pct_of_peak_sustained_elapsed']:.1f} | {m['sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | "
          f"{m['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']:.1f} | {m['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']:.1f} | "
          f"{m['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'] / cyc / 148:.3f} |")
