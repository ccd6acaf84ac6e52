set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_persistent.py -q -x 2>&1 | tail -5
[ "${PIPESTATUS[0]}" = "0" ] || { echo "persistent tests failed or hung: stop"; exit 1; }
for n in 100000 40000 10000; do
  timeout 60 python scripts/seq_rate.py $n 64 50 100 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/r2c10.log
  BANN_DEBUG_TCP=1 timeout 60 python scripts/seq_rate.py $n 2 50 100 2>&1 | grep "tcp\]" | tail -2 | tee -a gpurun_out/r2c10.log
done
