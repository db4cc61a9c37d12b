#!/bin/bash
# Round-2 measurement pass on one B200 (run through gpurun from the repo root):
#   1. the -m gpu parity suite, 2. the default bench line, 3. ncu launch list + one --set full capture of the per-batch
#   kernels of the same command (only after the plain run exited 0).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_gputest.log
tail -3 gpurun_out/${TAG}_gputest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
CMD="python bench.py --no-cpu-baseline --steps 2 --warmup 3"
export SMASH_NO_CHUNKS=1
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; exit 1; }
# launch list: the per-batch + tail kernels only (the index build alone is several thousand launches of its own)
KL='regex:^k_(mam|rec|sizes|emit|scan|publish|pair|bump|dd|varbin|perm|export|slot|sort|add_one)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KL" -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$?"
K='regex:^(k_mam_seed|k_mam_search|k_mam_verify|k_rec_build|k_rec_xe|k_sizes|k_emit_text|k_emit_copy)$'
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -s ${SKIP:-24} -c 8 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
tail -5 gpurun_out/${TAG}_ncu_full.log
# BASELINE.json configs[0]: the reference's CPU-runnable case (50 Mb, 4-byte index) next to 1 GPU
unset SMASH_NO_CHUNKS
python bench.py --workload config1 > gpurun_out/${TAG}_bench_config1.json 2> gpurun_out/${TAG}_bench_config1.err; echo "config1 rc=$?"
