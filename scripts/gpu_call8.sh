set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2c8_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2c8_gpu_tests.log; tail -12 gpurun_out/r2c8_gpu_tests.log
