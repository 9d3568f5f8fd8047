#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures (works without a GPU):

    python profiles/make_traffic_json.py 1m_deg3=gpurun_out/prof_1m_deg3.ncu-rep 6m_deg3_fp16=gpurun_out/prof_6m_deg3_fp16.ncu-rep [--source-note "..."]

Per workload: dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch of the roofline kernels
(bench.py names), averaged over the captured launches, tagged with bench.kernel_sources_sha() of the CURRENT sources -- run
it right after the capture, before touching a kernel.  bench.py reports `roofline.traffic` only when the tag still matches.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KERNELS = {"k_composite": "composite", "k_project_keys": "project", "k_emit_sorted": "emit_sorted"}


def to_bytes(v, u):
    return float(v) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u.lower(), 1)


def main():
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    prof = json.load(open(out_path)) if os.path.exists(out_path) else {}
    prof.setdefault("captures", {})
    note = sys.argv[sys.argv.index("--source-note") + 1] if "--source-note" in sys.argv else None
    for arg in sys.argv[1:]:
        if "=" not in arg or arg.startswith("--"):
            continue
        wl, rep = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        dram, inst, dur = {}, {}, {}
        for r in rows:
            name = r[idx["Kernel Name"]]
            key = next((v for k, v in KERNELS.items() if k in name), None)
            if key is None and "k_sort_onesweep" in name:
                # the roofline's sort kernel = a pass over the D instances (the largest grids' launches); depth passes separately
                key = "sort_onesweep" if to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) > 0 else None
            if key is None:
                continue
            b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            dram.setdefault(key, []).append(b)
            inst.setdefault(key, []).append(float(r[idx["smsp__inst_executed.sum"]]))
            du = units[idx["gpu__time_duration.sum"]].lower()
            dur.setdefault(key, []).append(float(r[idx["gpu__time_duration.sum"]]) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(du, 1.0))
        if "sort_onesweep" in dram:
            # keep the instance passes only: the launches that move the most bytes (a frame has 4 depth passes + 2 tile passes)
            order = sorted(range(len(dram["sort_onesweep"])), key=lambda i: -inst["sort_onesweep"][i])
            keep = order[: max(1, len(order) // 3)]
            for d in (dram, inst, dur):
                d["sort_onesweep"] = [d["sort_onesweep"][i] for i in keep]
        prof["captures"][wl] = {
            "kernel_sources_sha": bench.kernel_sources_sha(),
            "source": note or os.path.basename(rep),
            "dram_bytes_per_launch": {k: round(sum(v) / len(v)) for k, v in dram.items()},
            "warp_instructions_per_launch": {k: round(sum(v) / len(v)) for k, v in inst.items()},
            "ncu_duration_us": {k: round(sum(v) / len(v), 1) for k, v in dur.items()},
        }
    with open(out_path, "w") as f:
        json.dump(prof, f, indent=1)
    print(json.dumps(prof["captures"], indent=1))


if __name__ == "__main__":
    main()
