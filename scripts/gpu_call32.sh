set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c32_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -4 gpurun_out/r2c32_gpu_tests.log
rm -f gpurun_out/r2c32.jsonl
for w in cfg3r8 cfg3r8 cfg3 cfg3; do
timeout 600 python bench.py --workload $w --no-cpu-baseline --no-sequential >> gpurun_out/r2c32.jsonl 2> gpurun_out/r2c32.err; echo "$w exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c32.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:6], d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2))
PY
