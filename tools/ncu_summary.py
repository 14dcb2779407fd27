#!/usr/bin/env python
"""Reads .ncu-rep files (ncu must be on PATH) and prints a compact summary: duration, DRAM bytes, L2 sectors / hit rate,
instructions, IPC, occupancy, registers, top stall reasons.   python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        if len(r) == len(hdr):
            res.append({h: (v, u) for h, u, v in zip(hdr, units, r)})
    return res


def num(d, k):
    v, u = d.get(k, ("nan", ""))
    try:
        return float(v.replace(",", "")) * UNIT.get(u, 1)
    except ValueError:
        return float("nan")


def summary(d):
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): num(d, k)
              for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    top = sorted(stalls.items(), key=lambda x: -x[1])[:6]
    return {
        "kernel": d["Kernel Name"][0],
        "grid": d.get("launch__grid_size", ("", ""))[0], "block": d.get("launch__block_size", ("", ""))[0],
        "duration_ms": num(d, "gpu__time_duration.sum"),
        "dram_read_bytes": num(d, "dram__bytes_read.sum"), "dram_write_bytes": num(d, "dram__bytes_write.sum"),
        "dram_pct_of_peak": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l2_sectors": num(d, "lts__t_sectors.sum"), "l2_hit_pct": num(d, "lts__t_sector_hit_rate.pct"),
        "warp_instructions": num(d, "smsp__inst_executed.sum"), "ipc_active": num(d, "sm__inst_executed.avg.per_cycle_active"),
        "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": num(d, "launch__registers_per_thread"),
        "stall_cycles_per_issue_top": {k: round(v, 2) for k, v in top},
    }


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in load(p):
            print(json.dumps(summary(d)))
