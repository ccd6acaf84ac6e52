set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_cfg3_1gpu.json 2> gpurun_out/r2_bench_cfg3_1gpu.err; echo "bench exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_1gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['cpu_baseline']['value'],d['cpu_baseline']['sample_seconds']);print(d['sequential_schedule'])"
tail -3 gpurun_out/r2_bench_cfg3_1gpu.err
for w in cfg1 cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2_bench_${w}_1gpu.json 2> gpurun_out/r2_bench_${w}_1gpu.err; echo "bench $w exit $?"
  python -c "import json;d=json.loads(open('gpurun_out/r2_bench_${w}_1gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['sequential_schedule'])"
done
