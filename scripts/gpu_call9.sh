set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_persistent.py -q -x > gpurun_out/r2c9_persist.log 2>&1
echo "persistent tests exit $?" >> gpurun_out/r2c9_persist.log; tail -8 gpurun_out/r2c9_persist.log
for n in 100000 10000; do
  for p in 1 0; do
    HMC_PATH=$p timeout 300 python scripts/seq_rate.py $n 64 50 100 2>&1 | tail -1 | tee -a gpurun_out/r2c9_seq_rate.log
  done
done
BANN_DEBUG_TCP=1 timeout 300 python scripts/seq_rate.py 100000 2 50 100 2>&1 | grep "tcp\] us" | tail -1 | tee -a gpurun_out/r2c9_seq_rate.log
BANN_DEBUG_TCP=1 timeout 300 python scripts/seq_rate.py 10000 2 50 100 2>&1 | grep "tcp\] us" | tail -1 | tee -a gpurun_out/r2c9_seq_rate.log
