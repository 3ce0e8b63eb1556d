#!/usr/bin/env python
"""tools/ncu_summary.py — condense an .ncu-rep (read here, no GPU needed) into the counters the
north star asks for: issue-slot, FMA/ALU/XU pipe utilisation, divergence, shared-atomic traffic.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [events_per_launch] > profiles/xyz.md
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__warps_active.avg.per_cycle_active", "resident warps / SM"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / warp instruction (divergence)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots used, % of peak"),
    ("smsp__issue_active.avg.per_cycle_active", "warp instr issued / cycle / SMSP"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (FFMA/FFMA2/IMAD), % of peak"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe cycles active (IMAD.WIDE lives here)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (LOP3/ISETP/SHF/I2FP), % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe (MUFU), % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, % of peak"),
    ("sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "CBU (branch) pipe, % of peak"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe, % of peak"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared-memory atomic instructions (ATOMS)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "shared atomic wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared load wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe, % of peak"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (MUFU/shared)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
]


def main():
    rep = sys.argv[1]
    events = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"## launch {k}: `{d.get('Kernel Name', ('?',))[0]}`\n")
        print("| counter | value | unit | what |\n|---|---:|---|---|")
        for name, what in WANT:
            if name in d:
                print(f"| `{name}` | {d[name][0]} | {d[name][1]} | {what} |")
        if events:
            inst = float(d["smsp__inst_executed.sum"][0].replace(",", ""))
            lanes = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0])
            ns = float(d["gpu__time_duration.sum"][0].replace(",", ""))
            unit = d["gpu__time_duration.sum"][1]
            scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
            print(f"\nderived: {inst * 32 / events:.2f} warp-instruction slots per scatter event (lane-normalised), "
                  f"{inst * lanes / events:.2f} thread-instructions per event, "
                  f"{events / (ns * scale):.4g} events/s under the profiler (not a bench value)\n")


if __name__ == "__main__":
    main()
