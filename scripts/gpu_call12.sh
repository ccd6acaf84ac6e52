set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "release_byte_store or activations" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_bench_cfg3_1gpu.json 2> gpurun_out/r2_bench_cfg3_1gpu.err; echo "bench exit $?"; cat gpurun_out/r2_bench_cfg3_1gpu.json | cut -c1-1500
for w in cfg2 cfg4; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2_bench_${w}_1gpu.json 2> gpurun_out/r2_bench_${w}_1gpu.err; echo "bench $w exit $?"
  python -c "import json;d=json.loads(open('gpurun_out/r2_bench_${w}_1gpu.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['roofline'].get('kernel'),d['e2e']['value'])"
done
