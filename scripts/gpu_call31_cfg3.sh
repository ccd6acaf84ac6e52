set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
rm -f gpurun_out/r2c31.jsonl
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-sequential >> gpurun_out/r2c31.jsonl 2> gpurun_out/r2c31.err; echo "cfg3 exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c31.jsonl'):
    d = json.loads(l); print(d['config']['workload'][:6], d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],2))
PY
