#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the entries kept in profiles/r02_ncu_summary.json (per-kernel duration, grid,
registers, achieved warps, FP64-pipe and LSU utilisation, shared-memory conflicts, DRAM bytes per launch).
usage: ncu_summary.py report.ncu-rep "capture label" [summary.json to merge into]   (reads with `ncu -i ... --page raw --csv`)"""
import csv
import io
import json
import subprocess
import sys

rep, label = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "Ghz": 1.0, "Mhz": 1e-3}


def get(r, name, scaled=False):
    v = float(r[col[name]].replace(",", ""))
    return v * SCALE.get(units[col[name]], 1.0) if scaled else v


out = {}
for r in data:
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].replace("cmdr::", "")
    grid = r[col["launch__grid_size"]]
    key = name if name not in out else f"{name} grid {grid}"
    rd, wr = get(r, "dram__bytes_read.sum", True), get(r, "dram__bytes_write.sum", True)
    out[key] = {
        "capture": label,
        "duration_ms": get(r, "gpu__time_duration.sum", True),
        "grid_size": get(r, "launch__grid_size"), "block_size": get(r, "launch__block_size"),
        "registers_per_thread": get(r, "launch__registers_per_thread"),
        "warps_active_per_sm": get(r, "sm__warps_active.avg.per_cycle_active"),
        "warp_instructions": get(r, "smsp__inst_executed.sum"),
        "fp64_pipe_pct_of_peak": get(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lsu_data_pipe_wavefronts_pct": get(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "smem_bank_conflicts": get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": rd + wr,
        "sm_clock_ghz": get(r, "smsp__cycles_elapsed.avg.per_second", True),
    }
if len(sys.argv) > 3:
    try:
        merged = json.load(open(sys.argv[3]))
    except OSError:
        merged = {}
    merged.update(out)
    json.dump(merged, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
