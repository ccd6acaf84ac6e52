# experiment: tail parameters of k1_tc5 as warp-uniform values (uniform registers as FFMA2 operands) instead of shared-memory reads
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc5.py -x -q > gpurun_out/r2c35_tests.log 2>&1; echo "tc5 tests exit $?"; tail -2 gpurun_out/r2c35_tests.log
rm -f gpurun_out/r2c35.jsonl
for w in cfg3s cfg3s cfg3; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-sequential >> gpurun_out/r2c35.jsonl 2> gpurun_out/r2c35.err; echo "$w exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c35.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:6], d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4))
PY
