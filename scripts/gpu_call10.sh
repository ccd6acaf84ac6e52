set -u
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_persistent.py -q -x 2>&1 | tail -15
