#!/usr/bin/env python
"""Turns `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small tracked files under profiles/:
  profiles/r02_ncu_<tag>_raw.csv       the raw page (every metric of the captured launch)
  profiles/r02_ncu_summary.json        per workload: dram bytes, instructions, issue utilisation, ... (bench.py reads it)
Usage: python scripts/ncu_summary.py 2b=gpurun_out/r2_ncu_2b.ncu-rep 2a=gpurun_out/r2_ncu_2a.ncu-rep [build=...]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    "gpu__time_duration.sum": "duration_us_under_ncu",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "smsp__inst_executed.sum": "inst_executed",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "smsp__issue_active.avg.per_cycle_active": "issue_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.per_cycle_active": "warps_active",
    "launch__registers_per_thread": "registers",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct",
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
out = {}
for arg in sys.argv[1:]:
    tag, rep = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    dst = os.path.join(ROOT, "profiles", f"r02_ncu_{tag}_raw.csv")
    open(dst, "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None, "report": os.path.basename(rep)}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            x = float(v.replace(",", ""))
            if u in UNIT_SCALE:
                x *= UNIT_SCALE[u]
            d[WANT[h]] = x
        elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            d.setdefault("stalls_per_issue", {})[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v)
    if "dram_read_bytes" in d:
        d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
    if "stalls_per_issue" in d:
        d["stalls_per_issue"] = dict(sorted(d["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:6])
    out[tag] = d
path = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
prev = json.load(open(path)) if os.path.exists(path) else {}
prev.update(out)
json.dump(prev, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
