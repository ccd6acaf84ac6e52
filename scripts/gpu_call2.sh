# round 2, call 2: the whole GPU suite (grouped chain, activations, ...) + bench smoke
set -u
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2c2_gpu_tests.log 2>&1
echo "gpu tests exit $?" >> gpurun_out/r2c2_gpu_tests.log
tail -30 gpurun_out/r2c2_gpu_tests.log
