# round 2, k1_tc5 as the default: the whole GPU suite, A/B of the three variants on one box, cfg3 bench line, ncu capture
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c27_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/r2c27_gpu_tests.log
rm -f gpurun_out/r2c27_ab_cfg3s.jsonl
for v in four five-plain five four five-plain five; do
  timeout 300 python bench.py --workload cfg3s --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c27_ab_cfg3s.jsonl 2> gpurun_out/r2c27_ab.err; echo "cfg3s $v exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c27_ab_cfg3s.jsonl'):
    d = json.loads(l); print(d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4))
PY
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c27_cfg3.json 2> gpurun_out/r2c27_cfg3.err; echo "cfg3 exit $?"
python -c "import json;d=json.loads(open('gpurun_out/r2c27_cfg3.json').read().strip().splitlines()[-1]);print('cfg3', d['value'], d['k1_ms'], d['roofline']['frac'], d['e2e']['value'], d['roofline']['kernel'][:10], d['sequential_schedule'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_tc5 -s 4 -c 1 -o gpurun_out/r2c27_k1_tc5 \
  python bench.py --workload cfg3s --steps 3 --warmup 3 --no-cpu-baseline --no-sequential > gpurun_out/r2c27_ncu.log 2>&1
echo "ncu exit $?"
