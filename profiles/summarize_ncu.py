#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the per-kernel summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--traffic-json WORKLOAD]

Reads the report with `ncu -i REP --page raw --csv` (works without a GPU) and prints, per captured launch:
duration, DRAM bytes read/written (the `traffic` figure of bench.py's roofline), DRAM and SM throughput as % of
peak, issue rate, registers, occupancy limiters and the top warp-stall reasons.
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys

WANT = [
    ("dur_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("ipc_smsp", "smsp__issue_active.avg.per_cycle_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("regs", "launch__registers_per_thread"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("grid", "launch__grid_size"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]


def load(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value: str, unit: str) -> float:
    v = float(value)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep = sys.argv[1]
    hdr, units, rows = load(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    traffic = {}
    print(f"# ncu summary of {rep}\n")
    print("| kernel | " + " | ".join(k for k, _ in WANT) + " | top stalls (warps per issue) |")
    print("|---|" + "---|" * (len(WANT) + 1))
    for r in rows:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("splat::", "")
        cells = []
        for key, metric in WANT:
            if metric not in idx:
                cells.append("-")
                continue
            v, u = r[idx[metric]], units[idx[metric]]
            if key in ("dram_rd_MB", "dram_wr_MB"):
                cells.append(f"{to_bytes(v, u) / 1e6:.2f}")
            elif key == "dur_us":
                f = float(v)
                cells.append(f"{f / 1e3:.1f}" if u == "ns" else (f"{f:.1f}" if u in ("us", "usecond") else f"{f:.3f} {u}"))
            elif key in ("warp_inst", "grid", "regs"):
                cells.append(f"{float(v):.0f}")
            else:
                cells.append(f"{float(v):.2f}")
        top = sorted(((float(r[idx[h]]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls), reverse=True)[:4]
        print(f"| {name} | " + " | ".join(cells) + " | " + ", ".join(f"{n} {v:.1f}" for v, n in top) + " |")
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic.setdefault(name, []).append(rd + wr)
    if "--traffic-json" in sys.argv:
        wl = sys.argv[sys.argv.index("--traffic-json") + 1]
        print("\n```json")
        print(json.dumps({wl: {k: sum(v) / len(v) for k, v in traffic.items()}}, indent=1))
        print("```")


if __name__ == "__main__":
    main()
