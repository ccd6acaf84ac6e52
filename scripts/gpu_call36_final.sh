# end of round 2: smoke() and the default bench command (with the CPU baseline and the sequential-schedule key), reference arm
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c36_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2c36_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_cfg3_1gpu_final.json 2> gpurun_out/r2c36_bench.err; echo "bench exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2_bench_cfg3_1gpu_final.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel'][:12],d['e2e']['value'],d['cpu_baseline']['value'],d['gpu_launches'],d['clocks'],d['sequential_schedule']['visits_per_s'])"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 | cut -c1-400
