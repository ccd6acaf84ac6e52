# last call of round 2: the committed k1_tc5 (two slices of the deferred sums before the z0 wait): parity + smoke + one cfg3 line
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_tc5.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2c40_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r2c40_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c40_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2c40_smoke.log | cut -c1-200
timeout 60 python bench.py --no-cpu-baseline --no-sequential > gpurun_out/r2c40_cfg3.json 2> gpurun_out/r2c40.err; echo "bench exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2c40_cfg3.json').read().strip().splitlines()[-1]);print(d['value'],d['k1_ms'],d['roofline']['frac'],d['e2e']['value'])"
