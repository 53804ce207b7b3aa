#!/usr/bin/env python
"""Condenses one kernel of an `ncu --set full` report into the two files committed under profiles/:
<prefix>_ncu_full_details.csv (ncu --page details --csv of that kernel) and <prefix>_ncu_summary.json (the counters
DESIGN.md / bench.py quote + the stall-reason shares of the warp-state samples).

usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep <kernel id in the report> profiles/<prefix> [note]
"""
import csv
import io
import json
import subprocess
import sys

rep, kid, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""


def page(name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout


det = page("details")
rows = list(csv.reader(io.StringIO(det[det.index('"ID"'):])))
with open(prefix + "_ncu_full_details.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(rows[0])
    for r in rows[1:]:
        if r and r[0] == kid:
            w.writerow(r)

raw = page("raw")
rr = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
hdr, units = rr[0], rr[1]
row = [r for r in rr[2:] if r and r[0] == kid][0]
d = dict(zip(hdr, row))
u = dict(zip(hdr, units))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
out = {"_source": f"{rep} kernel id {kid}: {d.get('Kernel Name', '')[:80]}", "_note": note}
for k in KEYS:
    if k in d:
        out[k] = {"unit": u[k], "value": d[k]}
stall = {}
for k, v in d.items():
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            stall[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v.replace(",", ""))
        except ValueError:
            pass
tot = sum(stall.values()) or 1.0
out["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:10]}
with open(prefix + "_ncu_summary.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
