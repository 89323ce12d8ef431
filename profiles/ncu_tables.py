#!/usr/bin/env python
"""Turn an ncu report into the tables kept under profiles/:

    python profiles/ncu_tables.py gpurun_out/prof_r2.ncu-rep profiles/r2_ncu_full_table.md profiles/r2_ncu_dram_traffic.json "title" [cells]

Reads `ncu -i <rep> --page raw --csv` (or that CSV itself); one row per profiled launch."""
import csv
import io
import json
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time [us]", 1e-3), ("dram__bytes_read.sum", "dram rd [MB]", 1e-6), ("dram__bytes_write.sum", "dram wr [MB]", 1e-6),
        ("launch__registers_per_thread", "regs", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stalls / issue", 1)]


def to_float(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep, md, js, title = sys.argv[1:5]
    cells = int(sys.argv[5]) if len(sys.argv) > 5 else None
    if rep.endswith(".csv"):          # already exported on the GPU box (the .ncu-rep files are too large to bring back)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):] if '"ID"' in raw else raw
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}

    def unit_scale(name):      # ncu picks units per report (byte/Kbyte/Mbyte/Gbyte, us/ms/ns): normalise to bytes / ns
        u = units[idx[name]].lower() if name in idx else ""
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(u, 1.0)

    launches = []
    for r in body:
        name = r[idx["Kernel Name"]]
        grid = r[idx["Grid Size"]] if "Grid Size" in idx else r[idx.get("launch__grid_size", 0)]
        vals = {}
        for key, _, _ in COLS:
            vals[key] = to_float(r[idx[key]]) * (unit_scale(key) if ("bytes" in key or "time" in key) else 1.0) if key in idx else float("nan")
        launches.append((name, grid, vals))
    keep = list(range(len(launches)))
    with open(md, "w") as f:
        f.write(f"# {title}\n\nReplayed, cold-cache launches under `ncu --set full --clock-control none`: counters and shares, not bench times.  Source report: {rep}.\n\n")
        f.write("| # | kernel | grid | " + " | ".join(c[1] for c in COLS) + " |\n|---|---|---|" + "---|" * len(COLS) + "\n")
        for i in keep:
            name, grid, v = launches[i]
            short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            cellsv = []
            for key, _, _ in COLS:
                x = v[key]
                if "time" in key:
                    cellsv.append(f"{x * 1e-3:.1f}")
                elif "bytes" in key:
                    cellsv.append(f"{x * 1e-6:.1f}")
                else:
                    cellsv.append(f"{x:.4g}")
            f.write(f"| {i} | `{short}` | {grid} | " + " | ".join(cellsv) + " |\n")
    out = {"source": f"{rep}: {title}", "cells": cells, "kernels": {}}
    for i in keep:
        name, grid, v = launches[i]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        out["kernels"].setdefault(short, []).append(dict(launch=i, grid=grid, time_ms=v["gpu__time_duration.sum"] * 1e-6,
                                                          dram_bytes=v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]))
    with open(js, "w") as f:
        json.dump(out, f, indent=1)
    print(f"{len(keep)} launches -> {md}, {js}")


if __name__ == "__main__":
    main()
