# round 2, k1_tc5 variants 5 / 6 again with non-clobbering operand-image stores (the weight loads can be hoisted over them)
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc5.py -x -q > gpurun_out/r2c25_tests.log 2>&1; echo "tc5 tests exit $?"; tail -3 gpurun_out/r2c25_tests.log
rm -f gpurun_out/r2c25_ab_cfg3s.jsonl
for v in 4 5 6 4 5 6; do
  timeout 300 python bench.py --workload cfg3s --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c25_ab_cfg3s.jsonl 2> gpurun_out/r2c25_ab.err; echo "cfg3s $v exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c25_ab_cfg3s.jsonl'):
    d = json.loads(l); print(d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4))
PY
