set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
out=gpurun_out/r2_seq_rate.log; : > $out
echo "# scripts/seq_rate.py: sequential-exact schedule (bann_sweep, group size 1), one B200, 64 branches x 50 markers, widths [5,5,1], L = 100" >> $out
for n in 100000 10000; do
  echo "## N = $n, launch-per-step path (HMC_PATH=1)" >> $out
  HMC_PATH=1 timeout 60 python scripts/seq_rate.py $n 64 50 100 2>&1 | tail -1 >> $out
  echo "## N = $n, persistent kernel (default)" >> $out
  timeout 60 python scripts/seq_rate.py $n 64 50 100 2>&1 | tail -1 >> $out
  echo "## N = $n, persistent kernel, per-phase clocks of CTA 0 (BANN_DEBUG_TCP=1, 2 branches)" >> $out
  BANN_DEBUG_TCP=1 timeout 60 python scripts/seq_rate.py $n 2 50 100 2>&1 | grep "tcp\]" | tail -2 >> $out
done
cat $out | cut -c1-220
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r2c11_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2c11_gpu_tests.log; tail -6 gpurun_out/r2c11_gpu_tests.log
