#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report, one line each (reads `ncu --page raw --csv`).

    python profiles/ncu_brief.py gpurun_out/x.ncu-rep [--stalls kernel_regex]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
K = [("ms", "gpu__time_duration.sum"), ("dram_rd_GB", "dram__bytes_read.sum"), ("dram_wr_GB", "dram__bytes_write.sum"),
     ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
     ("inst_M", "smsp__inst_executed.sum"), ("thr/inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
     ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct")]
mult = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    out = []
    for lab, k in K:
        if k not in idx:
            continue
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]]
        if lab.endswith("_GB"):
            v *= mult.get(u, 1.0)
        if lab == "ms":
            v *= {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1)
        if lab == "inst_M":
            v /= 1e6
        out.append(f"{lab}={v:.3g}")
    st = sorted(((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]) for h, i in idx.items()
                 if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_issue_active.ratio")), reverse=True)[:4]
    print(f"{name:18s} " + " ".join(out) + "  stalls: " + ", ".join(f"{n} {v:.1f}" for v, n in st))
