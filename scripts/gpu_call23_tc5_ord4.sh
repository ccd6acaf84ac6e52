# round 2, k1_tc5 variant 5 (ORD 4: expansion under the second MUFU phase): parity + A/B on one box
set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc5.py -x -q > gpurun_out/r2c23_tests.log 2>&1; echo "tc5 tests exit $?"; tail -3 gpurun_out/r2c23_tests.log
rm -f gpurun_out/r2c23_ab_cfg3s.jsonl
for v in five 4 5 4 5; do
  timeout 300 python bench.py --workload cfg3s --k1-tc-variant $v --no-cpu-baseline --no-sequential >> gpurun_out/r2c23_ab_cfg3s.jsonl 2> gpurun_out/r2c23_ab.err; echo "cfg3s $v exit $?"
done
python - <<'PY'
import json
for l in open('gpurun_out/r2c23_ab_cfg3s.jsonl'):
    d = json.loads(l); print(d['roofline']['kernel'][:8], 'k1_ms', round(d['k1_ms'],4), 'value', round(d['value'],1), 'frac', round(d['roofline']['frac'],4))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_tc5 -s 4 -c 1 -o gpurun_out/r2c23_k1_tc5_v5 \
  python bench.py --workload cfg3s --k1-tc-variant 5 --steps 3 --warmup 3 --no-cpu-baseline --no-sequential > gpurun_out/r2c23_ncu.log 2>&1
echo "ncu exit $?"
