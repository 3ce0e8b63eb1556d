#!/usr/bin/env python
"""tools/ncu_facts.py — the few ncu counters bench.py's roofline block is built from, as JSON.

    python tools/ncu_facts.py <config> <report.ncu-rep> <events_per_launch> [profiles/r02_ncu_facts.json]

Reads one `ncu --set full` capture of the walk kernel (read here, no GPU needed) and merges
    {config: {kernel, block_threads, duration_ms, warp_instr, events, warp_instr_per_event,
              dispatch_slots_per_event, issue_active_pct, fma_heavy_pct, alu_pct, xu_pct, shared_pipe_pct,
              dram_bytes, sm_clock_ghz, report}}
into the facts file.  `dispatch_slots_per_event` counts instructions with 64-bit register operands
(IMAD.WIDE, FFMA2/FMUL2/FADD2) twice (profiles/r01_microbench_pipes.md), from the source page's
per-opcode executed counts.  bench.py multiplies its LIVE events/s by these per-event figures; nothing
measured under the profiler is ever reported as a throughput.
"""
import csv
import json
import re
import subprocess
import sys
from pathlib import Path


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v.replace(",", ""), u) for h, u, v in zip(hdr, units, vals)}


def wide_share(rep):
    """share of executed warp instructions that carry 64-bit register operands (two dispatch slots)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    col = ix.get("# Warp Instructions Executed", ix.get("Instructions Executed"))
    src = ix.get("Source")
    if col is None or src is None:
        return None
    total = wide = 0.0
    for r in rows[2:]:
        try:
            n = float(r[col])
        except (ValueError, IndexError):
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip()).split(" ")[0]
        total += n
        if op.startswith(("IMAD.WIDE", "FFMA2", "FMUL2", "FADD2")):
            wide += n
    return wide / total if total else None


def main():
    config, rep, events = sys.argv[1], sys.argv[2], float(sys.argv[3])
    path = Path(sys.argv[4] if len(sys.argv) > 4 else "profiles/r02_ncu_facts.json")
    d = raw_page(rep)

    def num(k):
        return float(d[k][0])

    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[d["gpu__time_duration.sum"][1]]
    byte_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum(num(k) * byte_scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    inst = num("smsp__inst_executed.sum")
    wide = wide_share(rep)
    fact = {
        "kernel": d["Kernel Name"][0],
        "block_threads": int(num("launch__block_size")),
        "registers": int(num("launch__registers_per_thread")),
        "duration_ms": num("gpu__time_duration.sum") * scale,
        "warp_instr": inst,
        "events": events,
        "warp_instr_per_event": inst * 32.0 / events,
        "dispatch_slots_per_event": inst * 32.0 / events * (1.0 + wide) if wide is not None else None,
        "issue_active_pct": num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "fma_heavy_pct": num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "alu_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "xu_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "shared_pipe_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "active_lanes_per_instr": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "dram_bytes": dram,
        "sm_clock_ghz": num("sm__cycles_elapsed.avg.per_second") * {"Ghz": 1.0, "Mhz": 1e-3, "hz": 1e-9}.get(d["sm__cycles_elapsed.avg.per_second"][1], 1.0),
        "report": Path(rep).name,
    }
    facts = json.loads(path.read_text()) if path.exists() else {}
    facts[config] = fact
    path.write_text(json.dumps(facts, indent=1, sort_keys=True) + "\n")
    print(json.dumps(fact, indent=1))


if __name__ == "__main__":
    main()
