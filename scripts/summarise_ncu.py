#!/usr/bin/env python
"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into tracked summaries under profiles/.
usage: python scripts/summarise_ncu.py <tag>"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
g = os.path.join(ROOT, "gpurun_out")
lines = ["# ncu summary `%s`" % tag, ""]

lpath = os.path.join(g, "launches_%s.csv" % tag)
if os.path.exists(lpath):
    shutil.copy(lpath, os.path.join(out_dir, "%s_launches.csv" % tag))
    rows = list(csv.reader(open(lpath)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            agg.setdefault(d["Kernel Name"], []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare SHARES)", "",
              "| kernel | launches | mean us | share |", "|---|---:|---:|---:|"]
    for n, v in agg.items():
        lines.append("| `%s` | %d | %.1f | %.1f%% |" % (n[:110], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    lines.append("")

rep = os.path.join(g, "prof_%s.ncu-rep" % tag)
rawcsv = os.path.join(g, "prof_%s_raw.csv" % tag)
if os.path.exists(rep) or os.path.exists(rawcsv):
    if os.path.exists(rawcsv):
        raw = open(rawcsv).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
            "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    ki = hdr.index("Kernel Name")
    lines += ["## `ncu --set full --clock-control none` (per launch)", "", "| metric | unit | " + " | ".join(
        "`%s`" % r[ki].replace("void ", "").split("(")[0][:40] for r in rows[2:]) + " |", "|---|---|" + "---:|" * len(rows[2:])]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            lines.append("| %s | %s | " % (w, units[i]) + " | ".join(r[i][:14] for r in rows[2:]) + " |")
    lines.append("")
    # dram traffic per launch of the forward kernels -> profiles/traffic.json (bench.py fills roofline.traffic from it)
    import json
    import re
    MODELS = {0: "FitzHughNagumo", 1: "LotkaVolterra", 2: "Lorenz", 3: "Prokaryote", 4: "JansenRit", 5: "OU2"}
    OPS = {0: "OP_DRAW", 1: "OP_RECOMPUTE", 2: "OP_LOGLIK", 3: "OP_INVSOLVE", 4: "OP_INVSOLVE_LL", 5: "OP_INIT", 6: "OP_SWEEP"}
    tpath = os.path.join(out_dir, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        mm = re.search(r"fwd_kernel<(?:dmt::)?Model<(\d+)>, *(\d+)", r[ki])
        if mm:
            key = "fwd_kernel<%s, %s>" % (MODELS[int(mm.group(1))], OPS[int(mm.group(2))])
            try:
                acc.setdefault(key, []).append(float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]])
            except ValueError:
                pass
    units_per_launch = int(os.environ.get("UNITS_PER_LAUNCH", "81920000"))
    for key, v in acc.items():
        v = [x for x in v if x == x]
        if not v:
            continue
        traffic[key] = {"dram_bytes_per_launch": int(sum(v) / len(v)), "units_per_launch": units_per_launch,
                        "source": "profiles/%s_summary.md (ncu --set full, mean of %d launches)" % (tag, len(v))}
    json.dump(traffic, open(tpath, "w"), indent=1)
open(os.path.join(out_dir, "%s_summary.md" % tag), "w").write("\n".join(lines))
print("\n".join(lines))
