# round 2, k1_tc5 (dedicated issuing warp): parity, A/B against k1_tc on one box, ncu capture of the new kernel
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc5.py -x -q > gpurun_out/r2c21_tests.log 2>&1; echo "tc5 tests exit $?"; tail -3 gpurun_out/r2c21_tests.log
for v in four five four five; do
  timeout 300 python bench.py --workload cfg3s --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c21_ab_cfg3s.jsonl 2> gpurun_out/r2c21_ab.err; echo "cfg3s $v exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c21_ab_cfg3s.jsonl'):
    d = json.loads(l); print(d['roofline']['kernel'][:8], 'k1_ms', d['k1_ms'], 'value', d['value'], 'frac', d['roofline']['frac'])
PY
timeout 600 python bench.py --k1-tc-variant five --no-cpu-baseline --no-sequential > gpurun_out/r2c21_cfg3_five.json 2> gpurun_out/r2c21_cfg3_five.err; echo "cfg3 five exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2c21_cfg3_five.json').read().strip().splitlines()[-1]);print('cfg3 five', d['value'], d['k1_ms'], d['roofline']['frac'], d['e2e']['value'])"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "five_warp or sampled" > gpurun_out/r2c21_fullsize.log 2>&1; echo "fullsize exit $?"; tail -3 gpurun_out/r2c21_fullsize.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_tc5 -s 4 -c 1 -o gpurun_out/r2c21_k1_tc5 \
  python bench.py --workload cfg3s --k1-tc-variant five --steps 3 --warmup 3 --no-cpu-baseline --no-sequential > gpurun_out/r2c21_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
