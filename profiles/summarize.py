#!/usr/bin/env python
"""Turn an ncu report (+ the launch-list CSV of the same command) into the committed summaries.

    python profiles/summarize.py gpurun_out/prof_r1_config2.ncu-rep gpurun_out/launches_r1_config2.csv r01 config2 "<command>"
"""
import collections
import csv
import json
import subprocess
import sys

rep, launches, rnd, wl, cmd = sys.argv[1:6]
raw = f"profiles/{rnd}_ncu_full_{wl}_raw.csv"
with open(raw, "w") as f:
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
subprocess.run(["cp", launches, f"profiles/{rnd}_launches_{wl}.csv"], check=True)
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
names = [r[idx["Kernel Name"]].split("(")[0] for r in rows[2:]]
mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for r, n in zip(rows[2:], names):
    # ncu prints one unit per column; a row's value is already expressed in it
    rd = float(r[idx["dram__bytes_read.sum"]]) * mult[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * mult[units[idx["dram__bytes_write.sum"]]]
    traffic[n] = {"dram_read_bytes": rd, "dram_write_bytes": wr}
h, agg = None, collections.OrderedDict()
for r in csv.reader(open(launches)):
    if len(r) > 5 and r[0] == "ID":
        h = r
        continue
    if h is None or len(r) != len(h):
        continue
    d = dict(zip(h, r))
    v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    agg.setdefault(d["Kernel Name"].split("(")[0].split("::")[-1], []).append(v)
with open(f"profiles/{rnd}_ncu_summary_{wl}.md", "w") as f:
    f.write(f"# {rnd} -- ncu evidence, workload {wl}\n\nCommand: `{cmd}`\n\n")
    f.write(f"* full capture (`--set full --clock-control none --import-source on`, one launch of each per-batch kernel of the timed step): raw export `{rnd}_ncu_full_{wl}_raw.csv`\n")
    f.write(f"* launch list of the same command (`--metrics gpu__time_duration.sum --clock-control none`): `{rnd}_launches_{wl}.csv`\n")
    f.write("* times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event stage times, not absolutes\n\n")
    f.write("| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(names) + "\n")
    for k in keys:
        if k in idx:
            f.write(f"| `{k}` [{units[idx[k]]}] | " + " | ".join(r[idx[k]] for r in rows[2:]) + " |\n")
    f.write("\n## launch list (mean device time per launch, 1 M reads per batch)\n\n| kernel | launches | mean ms |\n|---|---|---|\n")
    for k, v in agg.items():
        f.write(f"| `{k}` | {len(v)} | {sum(v) / len(v):.3f} |\n")
    per_batch = {k: sum(v) / len(v) for k, v in agg.items() if k in ("k_mam_seed", "k_mam_search", "k_mam_verify", "k_rec_build", "k_rec_xe", "k_sizes", "k_emit_text", "k_emit_copy", "k_pair_count", "k_pair_write")}
    tot = sum(per_batch.values())
    f.write("\nShare of the per-batch kernels: " + ", ".join(f"`{k}` {100 * v / tot:.0f} %" for k, v in per_batch.items()) + "\n")
# the capture is only evidence for the kernels it was taken from: bench.py compares this hash with the sources it runs
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_hash import kernel_source_hash
traffic["_kernel_sources_sha256"] = kernel_source_hash()
json.dump(traffic, open(f"profiles/{rnd}_dram_traffic_{wl}.json", "w"), indent=1)
print(open(f"profiles/{rnd}_ncu_summary_{wl}.md").read())
