P="import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']/1e6,1), round(d['ms_per_step'],3), {k:round(x,3) for k,x in d['stage_ms_per_step'].items()})"
for v in c6 c6s c7 c8; do
SMASH_B200_LIB=$PWD/smash_paper_b200/variants/libsmash_b200_$v.so python bench.py --no-cpu-baseline --workload config1 --steps 8 --warmup 3 2>/dev/null | python -c "$P" $v
done
