set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persistent.py -x -q -m gpu > gpurun_out/r2c7_persistent_tests.log 2>&1
echo "persistent tests exit $?" >> gpurun_out/r2c7_persistent_tests.log; tail -15 gpurun_out/r2c7_persistent_tests.log
for p in 1 0; do
  BANN_DEBUG_TCP=$p HMC_PATH=$p timeout 300 python scripts/seq_rate.py >> gpurun_out/r2c7_seq_rate.log 2>&1
  HMC_PATH=$p timeout 300 python scripts/seq_rate.py 10000 64 50 100 >> gpurun_out/r2c7_seq_rate.log 2>&1
done
cat gpurun_out/r2c7_seq_rate.log
