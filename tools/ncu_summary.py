#!/usr/bin/env python3
"""tools/ncu_summary.py -- condense ncu reports (gpurun_out/prof_*.ncu-rep) and the launch list
(gpurun_out/launches.csv) into profiles/<round>_ncu_summary.json + .md (run here, no GPU needed)."""
import csv
import glob
import json
import os
import subprocess
import sys
from collections import Counter, defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"

KEYS = {
    "duration_ms": "gpu__time_duration.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "warps_active_per_sm": "sm__warps_active.avg.per_cycle_active",
    "issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pipe_fp64_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "pipe_xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "pipe_lsu_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "inst_executed": "smsp__inst_executed.sum",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(uv):
    u, v = uv
    x = float(v.replace(",", ""))
    return x * UNIT_SCALE.get(u, 1.0)


summary = {"round": rnd, "kernels": {}}
for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
    m = raw(rep)
    name = os.path.basename(rep)[5:-8]
    k = {"kernel_name": m.get("Kernel Name", ("", ""))[1] if "Kernel Name" in m else name}
    for key, metric in KEYS.items():
        if metric in m:
            k[key] = num(m[metric])
    stalls = {h.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(uv[1])
              for h, uv in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    k["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    k["dram_traffic_bytes"] = k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0)
    summary["kernels"][name] = k

launches = os.path.join(OUT, "launches.csv")
if os.path.exists(launches):
    rows = list(csv.reader(open(launches)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(float)
    cnt = Counter()
    for r in rows:
        if len(r) == len(hdr) and r is not hdr and r[ix["Metric Name"]] == "gpu__time_duration.sum":
            nm = r[ix["Kernel Name"]].split("(")[0][:60]
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ix["Metric Unit"]], 1e-6)
            tot[nm] += float(r[ix["Metric Value"]].replace(",", "")) * scale
            cnt[nm] += 1
    total = sum(tot.values())
    summary["launch_list"] = {nm: {"launches": cnt[nm], "total_ms": tot[nm], "share": tot[nm] / total} for nm in sorted(tot, key=lambda n: -tot[n])}

os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
with open(os.path.join(REPO, "profiles", rnd + "_ncu_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
with open(os.path.join(REPO, "profiles", rnd + "_ncu_summary.md"), "w") as f:
    f.write("# ncu summary, %s (tools/profile_round.sh under gpurun, 1 x B200; --clock-control none)\n\n" % rnd)
    f.write("| kernel | ms | regs | warps/SM | issue % | FP64 % | ALU % | FMA % | XU % | LSU % | DRAM GB (r+w) | thr/inst | top stalls (per issue) |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for nm, k in summary["kernels"].items():
        f.write("| %s | %.2f | %d | %.1f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.2f | %.1f | %s |\n" % (
            nm, k.get("duration_ms", 0), k.get("registers_per_thread", 0), k.get("warps_active_per_sm", 0), k.get("issue_slots_busy_pct", 0),
            k.get("pipe_fp64_pct", 0), k.get("pipe_alu_pct", 0), k.get("pipe_fma_pct", 0), k.get("pipe_xu_pct", 0), k.get("pipe_lsu_pct", 0),
            k["dram_traffic_bytes"] / 1e9, k.get("threads_per_inst", 0),
            ", ".join("%s %.2f" % kv for kv in k["top_stalls_per_issue"].items())))
    if "launch_list" in summary:
        f.write("\nLaunch list of `bench.py --steps 2 --warmup 1 --em-pairs 2048` (10 000 pairs per step; cold-cache, serialised: compare shares):\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for nm, v in summary["launch_list"].items():
            f.write("| %s | %d | %.2f | %.3f |\n" % (nm, v["launches"], v["total_ms"], v["share"]))
print(json.dumps(summary)[:1500])
