#!/bin/bash
# 2-GPU check (gpurun --gpus 2): the multi-GPU tests that are skipped on one GPU, then the weak and strong bench lines.
set -u
mkdir -p gpurun_out
TAG=${1:-r02n2}
python -m pytest tests -m gpu -x -q -k "driver or shard or multigpu or bins_finish or comm" > gpurun_out/${TAG}_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_gputest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_weak.json 2> gpurun_out/${TAG}_bench_weak.err; echo "weak rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong --total-reads 16000000 > gpurun_out/${TAG}_bench_strong.json 2> gpurun_out/${TAG}_bench_strong.err; echo "strong rc=$?"
tail -c 400 gpurun_out/${TAG}_bench_weak.err; tail -c 400 gpurun_out/${TAG}_bench_strong.err
python - <<'P'
import json
for f in ("weak", "strong"):
    try:
        d = json.load(open(f"gpurun_out/r02n2_bench_{f}.json"))
        print(f, d["n_gpus"], d["scaling"], round(d["value"] / 1e6, 1), "M reads/s", round(d["ms_per_step"], 3), "ms/step e2e", round(d["e2e"]["value"] / 1e6, 1), "finish ms", round(d["tail_finish_ms"], 2))
    except Exception as e:
        print(f, "unreadable", e)
P
