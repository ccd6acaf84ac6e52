set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c28_gpu_tests.log 2>&1; echo "gpu tests exit $?"; tail -8 gpurun_out/r2c28_gpu_tests.log
