# fixed cost per CTA: the 8-GPU per-rank shard of cfg3 on one GPU (12.5k rows x 10 000 branches), before / after the prologue rework
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for v in four five five; do
timeout 300 python bench.py --workload cfg3r8 --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c29_cfg3r8.jsonl 2> gpurun_out/r2c29.err; echo "cfg3r8 exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c29_cfg3r8.jsonl'):
    d = json.loads(l); print(d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4))
PY
