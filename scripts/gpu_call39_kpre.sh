# experiment: two slices of the deferred sums before the wait for the super-tile's forward contraction (variant slot "five-plain")
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc5.py -x -q > gpurun_out/r2c39_tests.log 2>&1; echo "tc5 tests exit $?"; tail -2 gpurun_out/r2c39_tests.log
rm -f gpurun_out/r2c39.jsonl
for v in five five-plain five five-plain; do
timeout 600 python bench.py --workload cfg3 --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c39.jsonl 2> gpurun_out/r2c39.err; echo "$v exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c39.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:6], d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4))
PY
