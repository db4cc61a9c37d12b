set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02l_gputest.log
python profiles/ab_kernels.py --workload config2 --variants default,notailov@SMASH_NO_TAIL_OVERLAP=1 --steps 6 > gpurun_out/r02l_ab.jsonl 2> gpurun_out/r02l_ab.err; echo "ab rc=$?"
cat gpurun_out/r02l_ab.jsonl
timeout 600 python profiles/sweeps.py --workload config1 --mem --min-lens 20 --read-lens 150 --bins 1 > gpurun_out/r02l_sweeps_config1_mem.jsonl 2> gpurun_out/r02l_sweeps_config1_mem.err; echo "mem rc=$?"
cat gpurun_out/r02l_sweeps_config1_mem.jsonl | cut -c1-700
