set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for k in auto generic; do
timeout 300 python bench.py --workload cfg3s_321 --k1 $k --no-cpu-baseline --no-sequential > gpurun_out/r2_bench_cfg3s_321_$k.json 2> gpurun_out/r2_bench_cfg3s_321_$k.err; echo "bench $k exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3s_321_$k.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['kernel'],d['e2e']['value'])"
done
timeout 300 python bench.py --workload cfg3s --no-cpu-baseline --no-sequential 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('cfg3s',d['value'],d['ms_per_step'])"
