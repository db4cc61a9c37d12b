#!/bin/bash
# BASELINE configs[3]/[4] + MEM mode as measured lines (profiles/sweeps.py), then the config1 bench line (tests the
# reference-counters leg of the CPU baseline).  gpurun from the repo root.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 900 python profiles/sweeps.py --workload config2 > gpurun_out/${TAG}_sweeps_config2.jsonl 2> gpurun_out/${TAG}_sweeps_config2.err; echo "sweeps config2 rc=$?"
timeout 600 python profiles/sweeps.py --workload config1 --mem --min-lens 20 --read-lens 150 --bins 1 > gpurun_out/${TAG}_sweeps_config1_mem.jsonl 2> gpurun_out/${TAG}_sweeps_config1_mem.err; echo "sweeps config1+mem rc=$?"
timeout 600 python bench.py --workload config1 > gpurun_out/${TAG}_bench_config1.json 2> gpurun_out/${TAG}_bench_config1.err; echo "bench config1 rc=$?"
tail -c 1500 gpurun_out/${TAG}_sweeps_config2.err; tail -c 800 gpurun_out/${TAG}_sweeps_config1_mem.err; tail -c 1500 gpurun_out/${TAG}_bench_config1.err
