set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize_wide.py tests/test_cli_gpu.py -q -m gpu -x > gpurun_out/r2c5_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2c5_tests.log; tail -6 gpurun_out/r2c5_tests.log
timeout 600 python -m pytest tests/test_gpu_grouped_chain.py -q -m gpu -s -k r2_of > gpurun_out/r2c5_r2.log 2>&1
grep -E "^R2 |passed|failed" gpurun_out/r2c5_r2.log
timeout 600 python scripts/full_chain.py --individuals 20000 --test-individuals 5000 --branches 64 --iterations 10 --group-size 16 --causal-branches 16 > gpurun_out/r2c5_fullchain_small.log 2>&1
echo "full_chain exit $?"; tail -3 gpurun_out/r2c5_fullchain_small.log | cut -c1-1200
