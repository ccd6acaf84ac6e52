# k1_tc5 prologue rework: parity, then the 8-GPU shard shape (cfg3r8) and cfg3s
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc5.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2c30_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/r2c30_tests.log
rm -f gpurun_out/r2c30.jsonl
for w in cfg3r8 cfg3r8 cfg3s cfg3s; do
timeout 300 python bench.py --workload $w --no-cpu-baseline --no-sequential >> gpurun_out/r2c30.jsonl 2> gpurun_out/r2c30.err; echo "$w exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c30.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:6], d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4))
PY
